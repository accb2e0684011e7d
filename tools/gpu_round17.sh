#!/bin/bash
# Full GPU parity suite (no -x), step-time diagnosis (drift / host enqueue), C5 bench (F=512, 1024 rays).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.json gpurun_out/baselines.json
timeout -k 5 1500 python -m pytest tests -m gpu -q --timeout 240 > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -15
timeout 600 python tools/diag_step.py > gpurun_out/diag_step.txt 2>&1; cat gpurun_out/diag_step.txt | tail -16
timeout 600 python bench.py --feature-dim 512 --rays 1024 --width 648 --height 484 --frames 60 --render-frames 1 --no-cpu-baseline \
    > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_c5.json'))
    print('C5 ms/step', d['ms_per_step'], 'rays/s', d['value'], 'e2e', d['e2e']['value'], 'spr', d['config']['samples_per_ray'], d['config'].get('alive_samples_per_ray'))
    print(d.get('exact_compositing')); print(d.get('phases_ms')); print(d.get('render'))
except Exception as e:
    print('C5 bench failed', e)
PY
tail -3 gpurun_out/bench_c5.err
