#!/bin/bash
# Source-level ncu capture of the wide-head GEMM kernel inside a C5 training step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AL_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_gemm_tc -c 5 \
    -o gpurun_out/gemm_c5 -f python bench.py --feature-dim 512 --rays 1024 --width 648 --height 484 --frames 60 \
    --ncu-range 1 --pretrain 1500 --no-cpu-baseline > gpurun_out/ncu_gemm_c5.log 2>&1
tail -2 gpurun_out/ncu_gemm_c5.log; ls -la gpurun_out/gemm_c5.ncu-rep
