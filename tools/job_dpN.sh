#!/bin/bash
# N-GPU job: bench.py under torchrun on all visible GPUs + the exchange micro-benchmark.   gpurun --gpus N -- 'bash tools/job_dpN.sh TAG'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=$1
N=$(nvidia-smi -L | wc -l)
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --c5-steps 0 --no-early-leg > gpurun_out/bench_${TAG}_dp$N.out 2> gpurun_out/bench_${TAG}_dp$N.err
echo "dp$N exit $?"; grep -c "NCCL INFO" gpurun_out/bench_${TAG}_dp$N.out; grep -m2 -E "nranks|NVLS" gpurun_out/bench_${TAG}_dp$N.out | cut -c1-200
tail -1 gpurun_out/bench_${TAG}_dp$N.out > gpurun_out/bench_${TAG}_dp$N.json
python tools/show_bench.py gpurun_out/bench_${TAG}_dp$N.json | head -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    tools/bench_exchange.py > gpurun_out/exchange_${TAG}_${N}gpu.json 2> gpurun_out/exchange_${TAG}_${N}gpu.err
cat gpurun_out/exchange_${TAG}_${N}gpu.json; tail -2 gpurun_out/exchange_${TAG}_${N}gpu.err
