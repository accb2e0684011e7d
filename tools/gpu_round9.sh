#!/bin/bash
# Round-1 re-entry check: all GPU parity tests (incl. wide heads in the fused field), bench, smoke.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 1200 python -m pytest tests -m gpu -q --timeout 180 > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_r9.json 2> gpurun_out/bench_r9.err; cat gpurun_out/bench_r9.json; tail -3 gpurun_out/bench_r9.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
