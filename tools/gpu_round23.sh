#!/bin/bash
# 2 GPUs: bench with fresh batches on every step, peer vs NCCL gradient exchange, and the 1-GPU line; C5 with the new GEMM loop.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_mlp_wide_gpu.py tests/test_field_gpu.py tests/test_trainer_gpu.py -q --timeout 240 2>&1 | tail -4
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_dp1.json 2> gpurun_out/bench_dp1.err
for mode in peer nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py \
      --gpus 2 --steps 200 --warmup 20 --grad-exchange $mode --no-cpu-baseline \
      > gpurun_out/bench_dp2_$mode.json 2> gpurun_out/bench_dp2_$mode.err
done
python - <<'PY'
import json
for f in ('dp1','dp2_peer','dp2_nccl'):
    try:
        d=json.loads(open(f'gpurun_out/bench_{f}.json').read().strip().splitlines()[-1])
        print(f, 'ms/step', round(d['ms_per_step'],4), 'rays/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'repeat', round(d['value_repeat_after_e2e']['value']),
              'exact', round(d['exact_compositing']['value']), 'spr', d['config']['samples_per_ray'], d['config']['alive_samples_per_ray'], d['config']['grad_exchange'], 'render', d['render']['value'] if d.get('render') else None)
    except Exception as e:
        print(f, 'failed', e)
PY
wc -l gpurun_out/bench_dp2_peer.json
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
