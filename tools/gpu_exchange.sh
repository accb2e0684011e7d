#!/bin/bash
# N GPUs (gpurun --gpus N): gradient exchange + optimiser timed alone (NCCL all-reduce + Adam vs the peer-memory kernel),
# then the data-parallel bench in both forms.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    tools/bench_exchange.py > gpurun_out/exchange_${N}gpu.json 2> gpurun_out/exchange_${N}gpu.err
tail -1 gpurun_out/exchange_${N}gpu.json; tail -3 gpurun_out/exchange_${N}gpu.err
for mode in peer nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py \
      --gpus $N --steps 200 --warmup 20 --grad-exchange $mode --no-cpu-baseline --render-frames 2 \
      > gpurun_out/bench_dp${N}_$mode.json 2> gpurun_out/bench_dp${N}_$mode.err
  python - $mode $N <<'PY'
import json,sys
try:
    d=json.loads(open(f'gpurun_out/bench_dp{sys.argv[2]}_{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print(sys.argv[1], 'ms/step', round(d['ms_per_step'],4), 'rays/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'repeat', round(d['value_repeat_after_e2e']['value']),
          'spr', d['config']['samples_per_ray'], d['config']['alive_samples_per_ray'], d['config']['grad_exchange'], 'render', d['render']['value'] if d.get('render') else None)
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
  tail -2 gpurun_out/bench_dp${N}_$mode.err
done
