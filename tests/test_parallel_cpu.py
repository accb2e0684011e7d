"""Host-side logic of the ray-sharded data-parallel path on CPU: world_size-2 gloo process group,
gradient sum all-reduce in place on param.grad, mean folded into the optimiser's grad_scale, parameter
broadcast, frame sharding.  (The kernels themselves are covered by the -m gpu tests.)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from autolabel_b200 import parallel
    r, w, lr = parallel.init_distributed(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(rank)                       # replicas start different on purpose
    model = torch.nn.Linear(4, 3)
    model.register_buffer("grid", torch.full((5,), float(rank)))
    parallel.broadcast_parameters(model)
    ref = [p.detach().clone() for p in model.parameters()]

    class Opt:                                    # stands in for FusedAdam (only grad_scale matters here)
        grad_scale = 1.0
    opt = Opt()
    sync = parallel.GradientAllReduce(model.parameters(), opt)
    assert opt.grad_scale == 1.0 / world
    for i, p in enumerate(model.parameters()):
        p.grad = torch.full_like(p, float(rank + 1 + i))
    sync()
    sums = [float(p.grad.flatten()[0]) for p in model.parameters()]
    # a rank that produced no gradient for a tensor must still join the collective
    for p in model.parameters():
        p.grad = None if rank == 1 else torch.ones_like(p)
    sync()
    sums2 = [float(p.grad.flatten()[0]) for p in model.parameters()]
    out.put((rank, [t.tolist() for t in ref], model.grid.tolist(), sums, sums2,
             parallel.shard_frames(10, rank, world)))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gradient_allreduce_gloo_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (r0, ref0, grid0, sums0, sums0b, frames0), (r1, ref1, grid1, sums1, sums1b, frames1) = res
    assert ref0 == ref1 and grid0 == grid1 == [0.0] * 5            # broadcast from rank 0
    assert sums0 == sums1 == [1 + 2, 2 + 3]                        # SUM over ranks (mean is applied by the optimiser)
    assert sums0b == sums1b == [1.0, 1.0]
    assert frames0 == [0, 2, 4, 6, 8] and frames1 == [1, 3, 5, 7, 9]


def test_single_process_is_a_noop():
    from autolabel_b200 import parallel
    m = torch.nn.Linear(2, 2)
    sync = parallel.GradientAllReduce(m.parameters())
    for p in m.parameters():
        p.grad = torch.ones_like(p)
    sync()
    assert all(float(p.grad.sum()) == p.numel() for p in m.parameters())
    assert parallel.shard_frames(5, 0, 1) == [0, 1, 2, 3, 4]


def test_shard_bounds_cover_the_vector_once():
    """PeerShardedAdam's ownership map: aligned, disjoint, complete, for every world size the bench runs."""
    from autolabel_b200.parallel import shard_bounds
    for n in (14262480 + 62464, 1000, 4, 0, 62464):
        for world in (1, 2, 4, 8, 3):
            prev = 0
            for r in range(world):
                b, e = shard_bounds(n, r, world)
                assert b == prev and b <= e <= n and b % 4 == 0 and (e % 4 == 0 or e == n)
                prev = e
            assert prev == n
