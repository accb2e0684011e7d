#!/bin/bash
# New renderer tests (run() path, occupancy refresh) + the whole GPU suite.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.json gpurun_out/baselines.json
timeout 600 python -m pytest tests/test_renderer_gpu.py -q --timeout 240 2>&1 | tail -30
timeout -k 5 1500 python -m pytest tests -m gpu -q --timeout 240 > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -15
