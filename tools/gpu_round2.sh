#!/bin/bash
# tcgen05 MLP back end: parity first (short timeouts: a hung kernel must not hold the box), then A/B bench + ncu.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 180 python -m pytest tests/test_mlp_gpu.py -x -q --timeout 60 > gpurun_out/pytest_mlp_tc.log 2>&1
rc=$?; echo "mlp tc exit $rc" >> gpurun_out/pytest_mlp_tc.log; tail -15 gpurun_out/pytest_mlp_tc.log
if [ $rc -ne 0 ]; then
  # diagnose: forward only / backward only on one shape
  timeout -k 5 60 python tools/mlp_tc_debug.py > gpurun_out/mlp_tc_debug.log 2>&1; tail -40 gpurun_out/mlp_tc_debug.log
  exit 0
fi
timeout -k 5 900 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; cat gpurun_out/bench_tc.json
AL_MLP_BACKEND=mma timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_mma.json 2> gpurun_out/bench_mma.err; cut -c1-300 gpurun_out/bench_mma.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_tc.csv python bench.py --ncu-range 2 --pretrain 1000 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_mlp_bwd_tc|k_mlp_fwd_tc' -c 8 -o gpurun_out/prof_tc \
    python bench.py --ncu-range 1 --pretrain 1000 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
