#!/bin/bash
# CUDA-graph step: trainer tests, wide-field tests, bench with and without the graph.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 600 python -m pytest tests/test_trainer_gpu.py tests/test_field_gpu.py -m gpu -q --timeout 180 > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_r10_graph.json 2> gpurun_out/bench_r10.err; cat gpurun_out/bench_r10_graph.json; tail -3 gpurun_out/bench_r10.err
AL_NO_GRAPH=1 timeout 600 python bench.py --steps 200 --warmup 20 --render-frames 0 --no-cpu-baseline > gpurun_out/bench_r10_nograph.json 2>> gpurun_out/bench_r10.err; cut -c1-400 gpurun_out/bench_r10_nograph.json
