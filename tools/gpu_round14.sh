#!/bin/bash
# tests + short bench (training leg only) after a kernel change
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r14}
timeout -k 5 900 python -m pytest tests -m gpu -q --timeout 180 -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 200 --warmup 20 --render-frames 0 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; python - $TAG <<'PY'
import json, sys
d=json.load(open(f'gpurun_out/bench_{sys.argv[1]}.json'))
spr=d['config']['samples_per_ray']
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'spr', spr, 'ns/sample', d['ms_per_step']*1e6/(4096*spr))
print(d['phases_ms'])
print({k:round(v['ms'],4) for k,v in d['roofline']['all'].items()})
PY
tail -3 gpurun_out/bench_$TAG.err
