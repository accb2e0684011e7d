#!/bin/bash
# ncu --set full with source counters of the MLP kernels of one training step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
AL_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_mlp \
    -o gpurun_out/mlp_r12 -f python bench.py --ncu-range 1 --pretrain 1000 --no-cpu-baseline > gpurun_out/ncu_mlp.log 2>&1
tail -5 gpurun_out/ncu_mlp.log; ls -la gpurun_out/*.ncu-rep
