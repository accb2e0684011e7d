#!/usr/bin/env python
"""Gradient exchange + optimiser of the data-parallel step, timed alone on the bench model's parameters
(57 MB hash table + 62 k MLP values), one process per GPU:

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_exchange.py

  nccl : all_reduce(param.grad) x 5 tensors + the fused Adam on every rank  (parallel.GradientAllReduce + FusedAdam)
  peer : barrier + al_peer_adam_step + barrier                                (parallel.PeerShardedAdam),
         with multimem.ld_reduce / multimem.st when the fabric has multicast, and with plain peer loads / stores

CUDA events on the launching stream, max over ranks, rank 0 prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, dev, reps=50, warm=5):
    for _ in range(warm):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


def main():
    os.environ["NCCL_DEBUG"] = os.environ.get("AL_NCCL_DEBUG", "WARN")
    from autolabel_b200 import parallel
    from autolabel_b200.models import ALNetwork
    from autolabel_b200.trainer import configure_optimizer
    rank, world, local_rank = parallel.init_distributed()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)

    def model():
        m = ALNetwork(encoding='hg+freq', num_layers=2, hidden_dim=128, num_layers_color=2, hidden_dim_color=128,
                      hidden_dim_semantic=64, semantic_classes=2, bound=3.0, cuda_ray=True).to(dev)
        for p in m.parameters():
            if p.numel():
                p.grad = torch.randn_like(p) * 1e-3
        return m

    res = {"n_gpus": world}
    m = model()
    opt = configure_optimizer(m, lr=5e-3)
    sync = parallel.GradientAllReduce(m.parameters(), opt)

    def nccl_step():
        sync()
        opt.step()
    res["nccl_allreduce_plus_adam_ms"] = timed(nccl_step, dev)
    res["adam_only_ms"] = timed(opt.step, dev)
    del m, opt, sync
    for mc in (True, False):
        m = model()
        try:
            peer = parallel.PeerShardedAdam(m, lr=5e-3, use_multicast=mc)
            key = "peer_multicast_ms" if peer.multicast else "peer_loads_ms"
            if key not in res:
                res[key] = timed(peer.step, dev)
        except Exception as e:
            res["peer_error"] = repr(e)
        del m
    n_params = 14262480 + 62464
    res["bytes_exchanged_per_rank"] = {"nccl_allreduce": 2 * (world - 1) / world * n_params * 4,
                                       "peer": "loads: (W-1)/W of 57 MB read remotely + (W-1)/W of the parameters written remotely; "
                                               "multicast: one shard in (reduced in the switch), one shard out"}
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
