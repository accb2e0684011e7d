#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --steps 200 --warmup 20 --render-frames 0 --no-cpu-baseline > gpurun_out/bench_r11.json 2> gpurun_out/bench_r11.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r11.json'))
print(d['ms_per_step'], d['config']['samples_per_ray'], d['phases_ms'])
PY
tail -3 gpurun_out/bench_r11.err
