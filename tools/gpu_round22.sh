#!/bin/bash
# 2 GPUs: peer-memory gradient exchange + Adam (parity vs all-reduce + Adam), data-parallel bench in both modes.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_2gpu.txt 2>&1
timeout 600 python -m pytest tests/test_peer_adam_gpu.py -q -s --timeout 300 > gpurun_out/pytest_peer.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_peer.log
tail -25 gpurun_out/pytest_peer.log
timeout 600 python -m pytest tests/test_mlp_wide_gpu.py tests/test_field_gpu.py tests/test_mlp_gpu.py -q --timeout 240 2>&1 | tail -5
for mode in peer nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py \
      --gpus 2 --steps 200 --warmup 20 --grad-exchange $mode --no-cpu-baseline --render-frames 0 \
      > gpurun_out/bench_dp2_$mode.json 2> gpurun_out/bench_dp2_$mode.err
  python - $mode <<'PY'
import json,sys
try:
    d=json.load(open(f'gpurun_out/bench_dp2_{sys.argv[1]}.json'))
    print(sys.argv[1], 'ms/step', d['ms_per_step'], 'rays/s', d['value'], 'e2e', d['e2e']['value'], 'repeat', d['value_repeat_after_e2e']['value'], d['config']['grad_exchange'])
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
  tail -4 gpurun_out/bench_dp2_$mode.err
done
timeout 600 python bench.py --no-cpu-baseline --render-frames 0 > gpurun_out/bench_dp1.json 2> gpurun_out/bench_dp1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_dp1.json'))
print('1 GPU ms/step', d['ms_per_step'], 'rays/s', d['value'], 'e2e', d['e2e']['value'], 'repeat', d['value_repeat_after_e2e'])
PY
