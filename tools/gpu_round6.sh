#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q --timeout 120 -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
[ $rc -ne 0 ] && exit 0
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; cat gpurun_out/bench_a.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_a.csv python bench.py --ncu-range 2 --pretrain 1000 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_mlp_bwd_tc|k_mlp_fwd_tc|k_march_count|k_grid|k_encode' -c 14 -o gpurun_out/prof_a \
    python bench.py --ncu-range 1 --pretrain 1000 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
