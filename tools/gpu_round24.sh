#!/bin/bash
# GEMM path: tight window write-out, TMEM accumulation across wgrad splits.  Parity + C5 bench + C5 launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mlp_wide_gpu.py tests/test_field_gpu.py tests/test_trainer_gpu.py -q --timeout 240 2>&1 | tail -6
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
python tools/summarize_ncu.py launches gpurun_out/launches_c5.csv > gpurun_out/launches_c5.md 2>&1; head -8 gpurun_out/launches_c5.md
python - <<'PY'
import csv
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/launches_c5.csv') if l.startswith('"'))]
seq=[]
for r in rows:
    if r['Metric Name']!='gpu__time_duration.sum': continue
    n=r['Kernel Name']
    if 'gemm_tc' in n or 'wide_dout' in n or 'dgeo' in n:
        seq.append((n.split('(')[0][-10:], round(float(r['Metric Value'].replace(',',''))/1e3,1)))
print(seq[:19])
PY
tail -1 gpurun_out/launch_c5.log
