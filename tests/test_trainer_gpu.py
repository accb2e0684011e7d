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


@pytest.mark.parametrize("F,C", [(64, 2), (64, 40)])
def test_fused_step_equals_autograd_step(F, C):
    """C = 40: more classes than one 32-channel group (the loss kernel strides over the logits) and the wide
    semantic_out head."""
    from autolabel_b200.trainer import SimpleTrainer
    opt = SimpleNamespace(rgb_weight=1.0, depth_weight=0.1, semantic_weight=1.0, feature_weight=0.5, feature_loss=True, lr=5e-3)
    m1, data = _setup(F=F, C=C)
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
    assert abs(parts[1:].sum().item() - parts[0].item()) < 1e-5 * max(1.0, abs(parts[0].item()))   # fp32 atomic sums
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


@pytest.mark.parametrize("F,Fg,C", [(64, 64, 2), (64, 48, 5), (16, 0, 3), (64, 64, 40), (512, 512, 606)])
def test_loss_kernel_matches_pinned_port(F, Fg, C):
    """al_loss_fwd_bwd against oracle/run_path.loss_fn — the port that is pinned on the reference's own
    SimpleTrainer.train_step (tests/test_oracle_pinned.py, golden loss) — with all three masks active: depth only where
    gt > 0.01, cross-entropy only where label >= 0, features[:, :F_gt] (trainer.py:76-91).  Loss value and the gradients
    with respect to weights_sum / depth / every output channel, from autograd of the port on the same tensors."""
    from autolabel_b200._lib import call, ptr, stream_ptr
    from oracle import run_path
    N = 777
    K = 3 + C + F
    g = torch.Generator().manual_seed(F + 7 * C)
    ws = torch.rand(N, generator=g).cuda().requires_grad_(True)
    depth_raw = (torch.rand(N, generator=g) * 3).cuda().requires_grad_(True)
    out = torch.randn(N, K, generator=g).cuda().requires_grad_(True)
    norms = (torch.rand(N, generator=g) * 0.3 + 1.0).cuda()
    gt_rgb = torch.rand(N, 3, generator=g).cuda()
    gt_depth = (torch.rand(N, generator=g) * 3).cuda()
    gt_depth[torch.rand(N, generator=g).cuda() < 0.4] = 0.0          # invalid depth (sensor holes)
    gt_depth[5] = 0.01                                               # exactly on the threshold: excluded (strict >)
    gt_sem = torch.randint(-1, C, (N,), generator=g).cuda()
    gt_feat = torch.rand(N, Fg, generator=g).cuda() if Fg else None
    loss5, counts = torch.empty(5).cuda(), torch.empty(2, dtype=torch.int32).cuda()
    g_ws, g_d, g_o = torch.empty(N).cuda(), torch.empty(N).cuda(), torch.empty(N, K).cuda()
    call("al_loss_fwd_bwd", ptr(ws), ptr(depth_raw), ptr(out), N, C, F, ptr(norms), ptr(gt_rgb), ptr(gt_depth), ptr(gt_sem),
         ptr(gt_feat), Fg, 1.0, 0.1, 1.0, 0.5, 0.01, 1.0, ptr(loss5), ptr(counts), ptr(g_ws), ptr(g_d), ptr(g_o),
         stream_ptr(ws.device))
    # the port on the outputs the renderer epilogue forms from the same composited tensors (renderer.py:273-297)
    outputs = {'image': out[:, :3] + (1 - ws).unsqueeze(-1), 'depth': depth_raw / norms, 'semantic': out[:, 3:3 + C],
               'semantic_features': out[:, 3 + C:]}
    data = {'pixels': gt_rgb, 'depth': gt_depth, 'semantic': gt_sem}
    if Fg:
        data['features'] = gt_feat
    ref = run_path.loss_fn(outputs, data, rgb_weight=1.0, depth_weight=0.1, feature_weight=0.5, semantic_weight=1.0)
    ref.backward()
    assert int(counts[0]) == int((gt_depth > 0.01).sum()) and int(counts[1]) == int((gt_sem >= 0).sum())
    assert abs(loss5[0].item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
    assert torch.allclose(g_ws, ws.grad, atol=1e-8, rtol=1e-4)
    assert torch.allclose(g_d, depth_raw.grad, atol=1e-9, rtol=1e-4)
    assert torch.allclose(g_o, out.grad, atol=1e-8, rtol=1e-4)


def test_graph_step_contains_the_optimiser_and_takes_packed_batches():
    """Single GPU: the fused Adam (al_adam_multi, step count and learning rate in device memory) is part of the captured
    graph; a PackedBatch (one flat buffer, one copy) gives the same step as the dict it was packed from; a learning-rate
    change by a scheduler reaches the replayed graph."""
    from autolabel_b200.trainer import PackedBatch, SimpleTrainer
    opt = SimpleNamespace(rgb_weight=1.0, depth_weight=0.1, semantic_weight=1.0, feature_weight=0.5, feature_loss=True, lr=1e-3)
    m1, data = _setup()
    m2 = copy.deepcopy(m1)
    sched = lambda o: torch.optim.lr_scheduler.StepLR(o, gamma=0.5, step_size=1)
    t1 = SimpleTrainer('g', opt, m1, device='cuda:0', workspace=None, log_interval=0, update_interval=10 ** 9, use_graph=True,
                       lr_scheduler=sched)
    t2 = SimpleTrainer('e', opt, m2, device='cuda:0', workspace=None, log_interval=0, update_interval=10 ** 9, use_graph=False,
                       lr_scheduler=sched)
    packed_dev = PackedBatch.pack(data)
    packed_host = PackedBatch.pack(data, device='cpu', pin=True)
    assert packed_host.flat.is_pinned() and packed_dev.flat.is_cuda
    for k in data:
        assert torch.equal(packed_dev[k].reshape(-1), data[k].reshape(-1))
    for i in range(9):
        if i == 5:                                   # one StepLR epoch boundary: lr halves on both sides
            t1.lr_scheduler.step()
            t2.lr_scheduler.step()
        a = t1.train_one_step((data, packed_dev, packed_host)[i % 3]).item()
        b = t2.train_one_step(data).item()
        assert abs(a - b) < 2e-3 * max(1.0, abs(b)), (i, a, b)
    st = t1._graph_state
    assert st is not None and st['graph'] is not None and st['adam']
    assert t1.optimizer.param_groups[0]['lr'] == 5e-4
    for d in t1.optimizer._device_state():
        assert abs(float(d['lr'].item()) - 5e-4) < 1e-9
        assert int(d['step'].item()) == 9 == int(t1.optimizer.state[d['params'][0]]['step'])
    for (n1, p1), (n2, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        scale = max(p2.detach().abs().max().item(), 1e-6)
        assert (p1.detach() - p2.detach()).abs().max().item() < 5e-3 * scale, n1


def test_checkpoint_has_the_reference_key_set_and_resumes(tmp_path):
    """Checkpoint dictionary of the reference Trainer (torch_ngp/nerf/utils.py:1133-1163): keys, torch.optim.Adam-format
    optimiser state (loads into torch.optim.Adam itself), and a resumed trainer continues bit-for-bit like the original."""
    from autolabel_b200.trainer import SimpleTrainer
    opt = SimpleNamespace(rgb_weight=1.0, depth_weight=0.1, semantic_weight=1.0, feature_weight=0.5, feature_loss=True, lr=1e-3)
    m1, data = _setup()
    sched = lambda o: torch.optim.lr_scheduler.StepLR(o, gamma=0.5, step_size=2)
    t1 = SimpleTrainer('ngp', opt, m1, device='cuda:0', workspace=str(tmp_path), log_interval=0, update_interval=4,
                       use_graph=False, lr_scheduler=sched, ema_decay=0.95)
    for i in range(6):
        torch.manual_seed(i)
        t1.train_one_step(data)
    t1.ema.update()
    t1.lr_scheduler.step()
    t1.epoch = 1
    path = t1.save_checkpoint()
    state = torch.load(path, map_location='cpu', weights_only=False)
    assert {'epoch', 'global_step', 'stats', 'precision', 'mean_count', 'mean_density', 'opt0', 'lr_sched0', 'scaler', 'ema',
            'model'} <= set(state)
    assert set(state['ema']) >= {'decay', 'num_updates', 'shadow_params'}
    assert set(state['model']) == set(m1.state_dict())
    # the optimiser entry is torch.optim.Adam's own format
    m3 = copy.deepcopy(m1)
    ref_opt = torch.optim.Adam([{'params': list(m3.encoder.parameters())},
                                {'params': m3.network_parameters(), 'weight_decay': 1e-6}], lr=1e-3, betas=(0.9, 0.99), eps=1e-15)
    ref_opt.load_state_dict(state['opt0'])
    assert len(ref_opt.state) == len(t1.optimizer.state)
    # resume
    m2 = copy.deepcopy(m1)
    m2.mean_count, m2.local_step = 0, 0
    t2 = SimpleTrainer('ngp', opt, m2, device='cuda:0', workspace=str(tmp_path), log_interval=0, update_interval=4,
                       use_graph=False, lr_scheduler=sched, ema_decay=0.95)
    assert t2.global_step == 6 and t2.epoch == 1 and m2.mean_count == m1.mean_count
    assert t2.optimizer.param_groups[0]['lr'] == t1.optimizer.param_groups[0]['lr']
    m2.local_step = m1.local_step
    m2.step_counter.copy_(m1.step_counter)
    for i in range(3):
        torch.manual_seed(50 + i)
        a = t1.train_one_step(data).item()
        torch.manual_seed(50 + i)
        b = t2.train_one_step(data).item()
        assert abs(a - b) < 1e-3 * max(1.0, abs(b)), (i, a, b)
