"""SimpleTrainer on the GPU: the fused training step (one loss kernel, no autograd graph) against the reference-shaped
step (train_step of autolabel/trainer.py:54-94 through torch autograd) on the same model, batch and samples."""
import copy
from types import SimpleNamespace

import pytest
import torch

from tests.helpers import make_density_grid, make_rays

pytestmark = pytest.mark.gpu


def _setup(F=64, C=2, N=512, seed=0):
    from autolabel_b200 import raymarching as rm
    from autolabel_b200.models import ALNetwork
    torch.manual_seed(seed)
    m = ALNetwork(encoding='hg+freq', num_layers=2, hidden_dim=128, num_layers_color=2, hidden_dim_color=128,
                  hidden_dim_semantic=F, semantic_classes=C, bound=3.0, cuda_ray=True).cuda()
    with torch.no_grad():
        m._table().uniform_(-0.3, 0.3)
    grid = torch.from_numpy(make_density_grid(m.cascade, 128, seed=2, fill=0.04)).cuda()
    m.density_grid.copy_(grid)
    m.density_bitfield.copy_(rm.packbits(grid, 0.01))
    m.train()
    o, d = make_rays(N, 3.0, seed=3)
    g = torch.Generator().manual_seed(4)
    data = {'rays_o': torch.from_numpy(o).cuda(), 'rays_d': torch.from_numpy(d).cuda(),
            'direction_norms': (torch.rand(N, 1, generator=g) * 0.3 + 1.0).cuda(),
            'pixels': torch.rand(N, 3, generator=g).cuda(), 'depth': (torch.rand(N, generator=g) * 3 - 0.3).cuda(),
            'semantic': torch.randint(-1, C, (N,), generator=g).cuda(), 'features': torch.rand(N, F - 8, generator=g).cuda()}
    return m, data


def test_fused_step_equals_autograd_step():
    from autolabel_b200.trainer import SimpleTrainer
    opt = SimpleNamespace(rgb_weight=1.0, depth_weight=0.1, semantic_weight=1.0, feature_weight=0.5, feature_loss=True, lr=5e-3)
    m1, data = _setup()
    m2 = copy.deepcopy(m1)
    t1 = SimpleTrainer('a', opt, m1, device='cuda:0', workspace=None, log_interval=0, update_interval=10 ** 9, fused_step=True)
    t2 = SimpleTrainer('b', opt, m2, device='cuda:0', workspace=None, log_interval=0, update_interval=10 ** 9, fused_step=False)
    assert t1.fused_step_available() and not t2.fused_step_available()
    # gradients of one step (before the optimiser touches them)
    l1 = t1._fused_train_step(data)
    _, _, l2 = t2.train_step(data)
    l2.backward()
    assert abs(l1.item() - l2.item()) < 1e-5 * max(1.0, abs(l2.item()))
    parts = t1.last_loss_parts
    assert abs(parts[1:].sum().item() - parts[0].item()) < 1e-5
    for (n1, p1), (n2, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert n1 == n2
        if p2.grad is None:
            assert p1.grad is None or float(p1.grad.abs().max()) == 0, n1
            continue
        scale = max(p2.grad.abs().max().item(), 1e-12)
        assert (p1.grad - p2.grad).abs().max().item() < 2e-3 * scale, n1
    # and three full steps keep the two trainers together
    for t in (t1, t2):
        for p in t.model.parameters():
            if p.grad is not None:
                p.grad.zero_()
    for _ in range(3):
        a = t1.train_one_step(data)
        b = t2.train_one_step(data)
    assert abs(a.item() - b.item()) < 1e-3 * max(1.0, abs(b.item()))


def test_loss_kernel_edge_cases():
    """No labelled pixel, no valid depth, no feature target: the masked means fall back to 0 like the reference's
    clamp(min=1) denominators."""
    from autolabel_b200._lib import call, ptr, stream_ptr
    N, C, F = 100, 3, 16
    K = 3 + C + F
    g = torch.Generator().manual_seed(0)
    ws, depth, out = torch.rand(N, generator=g).cuda(), torch.rand(N, generator=g).cuda(), torch.randn(N, K, generator=g).cuda()
    norms, rgb = torch.ones(N).cuda(), torch.rand(N, 3, generator=g).cuda()
    gt_depth, gt_sem = torch.zeros(N).cuda(), torch.full((N,), -1, dtype=torch.long).cuda()
    loss5, counts = torch.empty(5).cuda(), torch.empty(2, dtype=torch.int32).cuda()
    g_ws, g_d, g_o = torch.empty(N).cuda(), torch.empty(N).cuda(), torch.empty(N, K).cuda()
    call("al_loss_fwd_bwd", ptr(ws), ptr(depth), ptr(out), N, C, F, ptr(norms), ptr(rgb), ptr(gt_depth), ptr(gt_sem), None, 0,
         1.0, 0.1, 1.0, 0.5, 0.01, 1.0, ptr(loss5), ptr(counts), ptr(g_ws), ptr(g_d), ptr(g_o), stream_ptr(ws.device))
    image = out[:, :3] + (1 - ws)[:, None]
    assert abs(loss5[0].item() - ((image - rgb) ** 2).mean().item()) < 1e-6
    assert loss5[2].item() == 0 and loss5[3].item() == 0 and loss5[4].item() == 0
    assert float(g_d.abs().max()) == 0 and float(g_o[:, 3:].abs().max()) == 0
    assert torch.allclose(g_o[:, :3], 2 * (image - rgb) / (3 * N), atol=1e-7)
    assert torch.allclose(g_ws, -(2 * (image - rgb) / (3 * N)).sum(1), atol=1e-7)


def test_graph_step_equals_eager_step():
    """The CUDA-graph replay of the fused step (one launch per step, re-captured when the sample budget changes)
    follows the kernel-by-kernel step: same losses, same per-step sample counters, host batches accepted."""
    from autolabel_b200.trainer import SimpleTrainer
    opt = SimpleNamespace(rgb_weight=1.0, depth_weight=0.1, semantic_weight=1.0, feature_weight=0.5, feature_loss=True, lr=1e-3)
    m1, data = _setup()
    m2 = copy.deepcopy(m1)
    # update_interval=4: the occupancy refresh changes mean_count -> the graph is re-captured several times
    t1 = SimpleTrainer('g', opt, m1, device='cuda:0', workspace=None, log_interval=0, update_interval=4, use_graph=True)
    t2 = SimpleTrainer('e', opt, m2, device='cuda:0', workspace=None, log_interval=0, update_interval=4, use_graph=False)
    host = {k: v.cpu().pin_memory() for k, v in data.items()}
    la, lb = [], []
    for i in range(14):
        torch.manual_seed(100 + i)                     # same occupancy-refresh noise on both sides
        la.append(t1.train_one_step(host if i % 2 else data).item())
        torch.manual_seed(100 + i)
        lb.append(t2.train_one_step(data).item())
    assert t1._graph_state is not None and t1._graph_state['graph'] is not None
    for a, b in zip(la, lb):
        assert abs(a - b) < 2e-3 * max(1.0, abs(b)), (la, lb)
    assert m1.mean_count > 0 and abs(m1.mean_count - m2.mean_count) <= 0.02 * m2.mean_count + 64
    c1, c2 = m1.step_counter.cpu(), m2.step_counter.cpu()
    assert (c1[:, 0] > 0).sum() == (c2[:, 0] > 0).sum()
    assert m1.local_step == m2.local_step
