#!/bin/bash
# Device-side sample budget (no re-capture per occupancy refresh), single-sync refresh, narrow composite backward:
# parity suite, step diagnosis, bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.json gpurun_out/baselines.json
timeout -k 5 1500 python -m pytest tests -m gpu -q --timeout 240 > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -15
timeout 600 python tools/diag_step.py > gpurun_out/diag_step.txt 2>&1; cat gpurun_out/diag_step.txt | tail -16
timeout 600 python bench.py > gpurun_out/bench_r18.json 2> gpurun_out/bench_r18.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r18.json'))
print('ms/step', d['ms_per_step'], 'rays/s', d['value'], 'e2e', d['e2e'], 'host', d['host_enqueue_ms_per_step'])
print('spr', d['config']['samples_per_ray'], d['config'].get('alive_samples_per_ray'), 'exact', d.get('exact_compositing'))
print(d['phases_ms']); print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['traffic'])
print(d.get('render'))
PY
tail -3 gpurun_out/bench_r18.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/baselines.json'))
for k,v in d['B2_reference_kernels_sm100a'].items():
    print(f"{k:38s} ref {v['ref_ms']:.4f}  ours {v['ours_ms']:.4f}  wrapper {v.get('ours_wrapper_ms',0):.4f}  speedup {v['speedup']:.2f}")
PY
