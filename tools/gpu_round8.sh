#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q --timeout 120 -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
[ $rc -ne 0 ] && exit 0
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; cat gpurun_out/bench_b.json; tail -3 gpurun_out/bench_b.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
