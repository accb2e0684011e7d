#!/bin/bash
# N-GPU data-parallel bench (ray-sharded training, gradient all-reduce over NCCL) + frame-sharded render leg.
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$N.txt
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err
echo "exit $?"; tail -5 gpurun_out/bench_dp$N.err; cat gpurun_out/bench_dp$N.json
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_dp1.json 2> gpurun_out/bench_dp1.err; cat gpurun_out/bench_dp1.json
