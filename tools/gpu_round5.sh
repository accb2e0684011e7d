#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q --timeout 120 -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
[ $rc -ne 0 ] && exit 0
AL_BWD_PARTS=2 timeout -k 5 600 python -m pytest tests/test_mlp_gpu.py tests/test_field_gpu.py -m gpu -q --timeout 120 -x -k "tcgen05" > gpurun_out/pytest_np2.log 2>&1; echo "np2 exit $?" >> gpurun_out/pytest_np2.log; tail -3 gpurun_out/pytest_np2.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_np4.json 2> gpurun_out/bench_np4.err; cat gpurun_out/bench_np4.json
AL_BWD_PARTS=2 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_np2.json 2> gpurun_out/bench_np2.err; cut -c1-200 gpurun_out/bench_np2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_tc.csv python bench.py --ncu-range 2 --pretrain 1000 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
AL_BWD_PARTS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_tc_np2.csv python bench.py --ncu-range 2 --pretrain 1000 --no-cpu-baseline > gpurun_out/ncu_launch2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_mlp_bwd_tc|k_mlp_fwd_tc|k_composite' -c 10 -o gpurun_out/prof_tc \
    python bench.py --ncu-range 1 --pretrain 1000 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
