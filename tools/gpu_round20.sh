#!/bin/bash
# Conflict-free tile loaders, staged GEMM window epilogue, live-row dout casts: parity, C2 bench, C5 bench + launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.json gpurun_out/baselines.json
timeout -k 5 1500 python -m pytest tests -m gpu -q --timeout 240 > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -15
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r20.json 2> gpurun_out/bench_r20.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r20.json'))
print('C2 ms/step', d['ms_per_step'], 'rays/s', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
print('spr', d['config']['samples_per_ray'], d['config'].get('alive_samples_per_ray'), 'exact', d.get('exact_compositing'))
print(d['phases_ms']); print({k:(round(v['ms'],4), round(v['frac'],3)) for k,v in d['roofline']['all'].items()})
print(d.get('render'))
PY
tail -3 gpurun_out/bench_r20.err
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
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_c5.csv python bench.py --feature-dim 512 --rays 1024 --width 648 --height 484 --frames 60 \
    --ncu-range 2 --pretrain 1500 --no-cpu-baseline > gpurun_out/launch_c5.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/launches_c5.csv > gpurun_out/launches_c5.md 2>&1; head -12 gpurun_out/launches_c5.md
