#!/bin/bash
# N-GPU job: bench.py under torchrun on all visible GPUs + the exchange micro-benchmark (skipped with a second argument
# `noexchange`).   gpurun --gpus N -- 'bash tools/job_dpN.sh TAG [noexchange]'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=$1
N=$(nvidia-smi -L | wc -l)
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --c5-steps 0 --no-early-leg > gpurun_out/bench_${TAG}_dp$N.out 2> gpurun_out/bench_${TAG}_dp$N.err
echo "dp$N exit $?"; grep -c "NCCL INFO" gpurun_out/bench_${TAG}_dp$N.out; grep -m2 -E "nranks|NVLS" gpurun_out/bench_${TAG}_dp$N.out | cut -c1-200
grep '^{"metric"' gpurun_out/bench_${TAG}_dp$N.out | tail -1 > gpurun_out/bench_${TAG}_dp$N.json; tail -1 gpurun_out/bench_${TAG}_dp$N.out | cut -c1-60
python tools/show_bench.py gpurun_out/bench_${TAG}_dp$N.json | head -3
[ "$2" = noexchange ] && exit 0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    tools/bench_exchange.py > gpurun_out/exchange_${TAG}_${N}gpu.json 2> gpurun_out/exchange_${TAG}_${N}gpu.err
cat gpurun_out/exchange_${TAG}_${N}gpu.json; tail -2 gpurun_out/exchange_${TAG}_${N}gpu.err
