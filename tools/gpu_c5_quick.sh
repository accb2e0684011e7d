timeout 300 python -m pytest tests/test_mlp_wide_gpu.py tests/test_field_gpu.py -q --timeout 240 2>&1 | tail -2
timeout 300 python bench.py --feature-dim 512 --rays 1024 --width 648 --height 484 --frames 60 --render-frames 1 --no-cpu-baseline --pretrain 2000 --steps 100 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c5.json"))
print(d["ms_per_step"],d["value"],d["config"]["samples_per_ray"],d["config"]["alive_samples_per_ray"])
p=d["phases_ms"];print(p["field_forward"],p["field_backward"],p["live_samples"],p["marched_samples"])
print(d["render"]["value"],d["render"]["samples_per_ray"])
PY
