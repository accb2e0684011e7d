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


def _state_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from autolabel_b200 import parallel
    parallel.init_distributed(backend="gloo")
    shapes = [(10, 2), (12,), (4, 4)]                     # 20 + 12 + 16 = 48 values, shards of 24
    n = 48
    b, e = parallel.shard_bounds(n, rank, world)
    full_m = torch.arange(n, dtype=torch.float32)
    full_v = torch.arange(n, dtype=torch.float32) * 0.5
    m = parallel.gather_shards(full_m[b:e].clone(), n, rank, world)
    v = parallel.gather_shards(full_v[b:e].clone(), n, rank, world)
    groups = [{'name': 'encoding', 'params': [0], 'lr': 1e-3, 'weight_decay': 0.0},
              {'name': 'net', 'params': [0, 0], 'lr': 1e-3, 'weight_decay': 1e-6}]
    sd = parallel.adam_state_dict_from_flat(m, v, 7, shapes, groups)
    out.put((rank, m.tolist(), v.tolist(), sd['state'][1]['exp_avg'].tolist(), [g['params'] for g in sd['param_groups']]))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_optimizer_state_gathers_into_torch_adam_format_gloo_world2():
    """PeerShardedAdam.state_dict(): every rank's moment shard is all-gathered and cut per parameter into
    torch.optim.Adam's state_dict layout, so a checkpoint written from N ranks resumes at any world size."""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_state_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, m, v, second, groups in res:
        assert m == [float(i) for i in range(48)] and v == [0.5 * i for i in range(48)]
        assert second == [float(i) for i in range(20, 32)]
        assert groups == [[0], [1, 2]]


def test_adam_state_round_trip_and_torch_adam_accepts_it():
    from autolabel_b200 import parallel
    ps = [torch.nn.Parameter(torch.randn(10, 2)), torch.nn.Parameter(torch.randn(12)), torch.nn.Parameter(torch.randn(4, 4))]
    opt = torch.optim.Adam([{'params': ps[:1]}, {'params': ps[1:], 'weight_decay': 1e-6}], lr=5e-3, betas=(0.9, 0.99), eps=1e-15)
    for p in ps:
        p.grad = torch.randn_like(p)
    opt.step()
    sd = opt.state_dict()
    m, v, step = parallel.flat_from_adam_state_dict(sd, 48, 'cpu')
    assert step == 1 and m.numel() == 48
    sd2 = parallel.adam_state_dict_from_flat(m, v, step, [tuple(p.shape) for p in ps], opt.param_groups)
    opt2 = torch.optim.Adam([{'params': ps[:1]}, {'params': ps[1:], 'weight_decay': 1e-6}], lr=5e-3, betas=(0.9, 0.99), eps=1e-15)
    opt2.load_state_dict(sd2)
    for p in ps:
        assert torch.equal(opt2.state[p]['exp_avg'], opt.state[p]['exp_avg'])
        assert torch.equal(opt2.state[p]['exp_avg_sq'], opt.state[p]['exp_avg_sq'])
    # every rank of a 4-way job finds its shard in the same file
    for r in range(4):
        b, e = parallel.shard_bounds(48, r, 4)
        assert torch.equal(m[b:e], torch.cat([opt.state[p]['exp_avg'].reshape(-1) for p in ps])[b:e])


def test_packed_batch_layout_round_trip():
    """One flat buffer per training batch (single H2D / D2D copy per step): views alias the buffer, offsets aligned."""
    from autolabel_b200.trainer import PackedBatch, batch_layout
    lay, nbytes = batch_layout(4096, 64)
    assert nbytes == 4096 * (3 + 3 + 1 + 3 + 1 + 64) * 4 + 4096 * 8 and lay['semantic'][0] % 8 == 0
    g = torch.Generator().manual_seed(0)
    d = {'rays_o': torch.rand(512, 3, generator=g), 'rays_d': torch.rand(512, 3, generator=g),
         'direction_norms': torch.rand(512, 1, generator=g), 'pixels': torch.rand(512, 3, generator=g),
         'depth': torch.rand(512, generator=g), 'semantic': torch.randint(-1, 2, (512,), generator=g),
         'features': torch.rand(512, 24, generator=g)}
    p = PackedBatch.pack(d)
    assert p.shape_key == (512, 24) and set(p) == set(d)
    for k in d:
        assert torch.equal(p[k].reshape(-1), d[k].reshape(-1)) and p[k].dtype == d[k].dtype
    q = PackedBatch(512, 24)
    q.flat.copy_(p.flat)                                     # ONE copy moves the whole batch
    assert all(torch.equal(q[k], p[k]) for k in d)
    assert 'features' not in PackedBatch(64, 0)
