#!/bin/bash
# Round-1 closing evidence (final state) on one B200: GPU parity tests + measured baselines (B1/B2), the full bench line, the CPU
# reference arm, the ncu launch list of the bench command, one `ncu --set full` capture of a whole training step
# (DRAM traffic per kernel -> roofline.traffic) and smoke().
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.json gpurun_out/baselines.json
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu_r1h.txt
timeout -k 5 1100 python -m pytest tests -m gpu -q --timeout 240 > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1h.log 2>&1; tail -1 gpurun_out/smoke_r1h.log
timeout 600 python bench.py > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1h.json'))
spr=d['config']['samples_per_ray']
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'spr', spr, 'alive', d['config'].get('alive_samples_per_ray'))
print('exact', d.get('exact_compositing'))
print(d['phases_ms'])
print({k:(round(v['ms'],4), round(v['frac'],3)) for k,v in d['roofline']['all'].items()})
print(d.get('render')); print(d.get('cpu_baseline')); print(d.get('clocks'))
PY
tail -3 gpurun_out/bench_r1h.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1h_reference.json 2>> gpurun_out/bench_r1h.err; cat gpurun_out/bench_r1h_reference.json
# launch list (graph replay: ncu profiles the kernel nodes)
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_r1h.csv python bench.py --ncu-range 2 --no-cpu-baseline > gpurun_out/launch_r1h.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/launches_r1h.csv > gpurun_out/launches_r1h.md 2>&1; head -34 gpurun_out/launches_r1h.md
# full capture of one step, kernel by kernel (no graph)
AL_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -o gpurun_out/step_r1h -f python bench.py --ncu-range 1 --no-cpu-baseline > gpurun_out/ncu_full_r1h.log 2>&1
tail -2 gpurun_out/ncu_full_r1h.log; ls -la gpurun_out/*.ncu-rep
python tools/summarize_ncu.py full gpurun_out/step_r1h.ncu-rep > gpurun_out/ncu_full_r1h.csv 2>gpurun_out/ncu_full_r1h.err; head -3 gpurun_out/ncu_full_r1h.csv
python tools/summarize_ncu.py traffic gpurun_out/step_r1h.ncu-rep > gpurun_out/roofline_traffic.json 2>/dev/null
grep -h ncu_range_steps gpurun_out/ncu_full_r1h.log gpurun_out/launch_r1h.log
timeout 600 python bench.py --feature-dim 512 --rays 1024 --width 648 --height 484 --frames 60 --render-frames 1 --no-cpu-baseline \
    > gpurun_out/bench_r1h_c5.json 2> gpurun_out/bench_r1h_c5.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1h_c5.json'))
print('C5 ms/step', d['ms_per_step'], 'rays/s', d['value'], 'e2e', d['e2e']['value'], 'exact', d['exact_compositing']['value'], 'render', d['render']['value'])
PY
