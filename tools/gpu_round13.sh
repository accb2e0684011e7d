#!/bin/bash
# Round-1 re-entry check: GPU parity tests, full bench line, launch list, and a source-level ncu capture of the
# sigma MLP forward/backward kernels (stall reasons per source line).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q --timeout 180 -x > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_r13.json 2> gpurun_out/bench_r13.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r13.json'))
spr=d['config']['samples_per_ray']
print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'spr', spr, 'ns/sample', d['ms_per_step']*1e6/(4096*spr))
print(d['phases_ms'])
print({k:round(v['ms'],4) for k,v in d['roofline']['all'].items()})
print(d.get('render'))
PY
tail -3 gpurun_out/bench_r13.err
# launch list (graph replay: ncu profiles the kernel nodes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_r13.csv python bench.py --ncu-range 2 --pretrain 1000 --no-cpu-baseline > gpurun_out/launch_r13.log 2>&1
python tools/summarize_ncu.py gpurun_out/launches_r13.csv > gpurun_out/launches_r13.md 2>&1; head -30 gpurun_out/launches_r13.md
# source-level capture of the sigma MLP kernels
AL_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    --kernel-name-base demangled -k 'regex:k_mlp_(fwd|bwd)_tc<48' \
    -o gpurun_out/mlp_r13 -f python bench.py --ncu-range 1 --pretrain 1000 --no-cpu-baseline > gpurun_out/ncu_mlp_r13.log 2>&1
tail -3 gpurun_out/ncu_mlp_r13.log; ls -la gpurun_out/*.ncu-rep
for i in 0 1; do
  ncu -i gpurun_out/mlp_r13.ncu-rep --page source --csv --print-source cuda,sass --launch-skip $i --launch-count 1 > gpurun_out/mlp_src_$i.csv 2>/dev/null
  python tools/ncu_source_hot.py gpurun_out/mlp_src_$i.csv 45 > gpurun_out/mlp_hot_$i.txt 2>&1
  rm -f gpurun_out/mlp_src_$i.csv
done
ncu -i gpurun_out/mlp_r13.ncu-rep --page raw --csv > gpurun_out/mlp_r13_raw.csv 2>/dev/null
head -50 gpurun_out/mlp_hot_0.txt gpurun_out/mlp_hot_1.txt
