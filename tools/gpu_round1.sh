#!/bin/bash
# One gpurun call: UMMA probe, GPU parity tests, bench at both occupancy thresholds, ncu launch list + full capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1
tools/run_probe.sh > /dev/null 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_t10.json 2> gpurun_out/bench_t10.err
timeout 600 python bench.py --steps 50 --warmup 10 --density-thresh 0.01 --no-cpu-baseline > gpurun_out/bench_t001.json 2> gpurun_out/bench_t001.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_t10.csv python bench.py --ncu-range 2 --pretrain 1000 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'k_mlp_bwd|k_mlp_fwd|k_encode_position|k_grid_bwd|k_composite' -c 16 -o gpurun_out/prof_r1 \
    python bench.py --ncu-range 1 --pretrain 1000 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_t10.json; cat gpurun_out/bench_t001.json | cut -c1-600
