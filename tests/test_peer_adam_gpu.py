"""Data-parallel gradient exchange + Adam as one peer-memory kernel (csrc/peer.cu, parallel.PeerShardedAdam) against
the path it replaces: sum of the replicas' gradients -> al_adam_step with grad_scale = 1 / world on every rank.
Needs two GPUs with peer access (skipped otherwise); run on the GPU box with `gpurun --gpus 2`."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, use_multicast, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from autolabel_b200 import parallel
    from autolabel_b200._lib import call, ptr, stream_ptr
    from autolabel_b200.models import ALNetwork
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(0)
    m = ALNetwork(encoding='hg+freq', num_layers=2, hidden_dim=128, num_layers_color=2, hidden_dim_color=128,
                  hidden_dim_semantic=64, semantic_classes=2, bound=3.0, cuda_ray=True).to(dev)
    with torch.no_grad():
        m._table().uniform_(-0.3, 0.3)
    parallel.broadcast_parameters(m)
    peer = parallel.PeerShardedAdam(m, lr=5e-3, use_multicast=use_multicast)
    params = peer.params
    ref_p = [p.detach().clone() for p in params]
    ref_m = [torch.zeros_like(p) for p in params]
    ref_v = [torch.zeros_like(p) for p in params]
    n_enc = len([q for q in m.encoder.parameters() if q.numel() > 0])
    worst = 0.0
    for step in range(1, 4):
        grads = []                                    # every rank can rebuild every rank's gradient
        for r in range(world):
            g = torch.Generator(device=dev).manual_seed(1000 * step + r)
            grads.append([torch.randn(p.shape, generator=g, device=dev) * 1e-2 for p in params])
        for p, gr in zip(params, grads[rank]):
            p.grad.copy_(gr)
        torch.cuda.synchronize()
        peer.step()
        torch.cuda.synchronize()
        for i, p in enumerate(params):                # the path it replaces: sum in rank order, then the fused Adam
            total = grads[0][i].clone()
            for r in range(1, world):
                total += grads[r][i]
            wd = 0.0 if i < n_enc else 1e-6
            call("al_adam_step", ptr(ref_p[i]), ptr(total), ptr(ref_m[i]), ptr(ref_v[i]), ref_p[i].numel(), 5e-3, 0.9, 0.99,
                 1e-15, wd, step, 1.0 / world, 1, stream_ptr(dev))
            err = (p.detach() - ref_p[i]).abs().max().item()
            worst = max(worst, err)
            assert err <= 1e-7, (step, i, err)
            assert float(p.grad.abs().max()) == 0.0, "step() leaves the gradient buffer cleared"
    # replicas identical
    flat = peer.flat_param.detach()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    for r in range(1, world):
        assert torch.equal(gathered[0], gathered[r])
    if rank == 0:
        out.put({"multicast": peer.multicast, "worst": worst, "n": peer.n, "shard": (peer.begin, peer.end)})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("use_multicast", [False, True])
def test_peer_sharded_adam_matches_allreduce_adam(use_multicast):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    mp.spawn(_worker, args=(2, _free_port(), use_multicast, out), nprocs=2, join=True)
    res = out.get()
    print(res)
    assert res["worst"] <= 1e-7
