#!/bin/bash
# Re-entry check: GPU parity tests + measured baselines (B1/B2), full bench line, launch list of the bench command,
# one `ncu --set full` capture of a whole training step (DRAM traffic per kernel -> roofline.traffic) and a source-level
# capture of the sigma MLP kernels (stall regions).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.json gpurun_out/baselines.json
timeout -k 5 1200 python -m pytest tests -m gpu -q --timeout 240 -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_r15.json 2> gpurun_out/bench_r15.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r15.json'))
spr=d['config']['samples_per_ray']
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'spr', spr, 'ns/sample', d['ms_per_step']*1e6/(4096*spr))
print(d['phases_ms'])
print({k:round(v['ms'],4) for k,v in d['roofline']['all'].items()})
print(d.get('render'))
PY
tail -3 gpurun_out/bench_r15.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r15_reference.json 2>> gpurun_out/bench_r15.err; cat gpurun_out/bench_r15_reference.json
# launch list (graph replay: ncu profiles the kernel nodes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_r15.csv python bench.py --ncu-range 2 --pretrain 1000 --no-cpu-baseline > gpurun_out/launch_r15.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/launches_r15.csv > gpurun_out/launches_r15.md 2>&1; head -32 gpurun_out/launches_r15.md
# full capture of one step, kernel by kernel (no graph), with source for the MLP kernels
AL_NO_GRAPH=1 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -o gpurun_out/step_r15 -f python bench.py --ncu-range 1 --pretrain 1000 --no-cpu-baseline > gpurun_out/ncu_full_r15.log 2>&1
tail -2 gpurun_out/ncu_full_r15.log; ls -la gpurun_out/*.ncu-rep
python tools/summarize_ncu.py full gpurun_out/step_r15.ncu-rep > gpurun_out/ncu_full_r15.csv 2>gpurun_out/ncu_full_r15.err; head -5 gpurun_out/ncu_full_r15.csv
ncu -i gpurun_out/step_r15.ncu-rep --page source --csv --print-source sass --kernel-name-base demangled -k 'regex:k_mlp_(fwd|bwd)_tc<48' > gpurun_out/mlp_src_r15.csv 2>/dev/null
for i in 0 1; do python tools/ncu_regions.py gpurun_out/mlp_src_r15.csv $i 1.0 > gpurun_out/mlp_regions_$i.txt 2>&1; done
head -60 gpurun_out/mlp_regions_0.txt gpurun_out/mlp_regions_1.txt
