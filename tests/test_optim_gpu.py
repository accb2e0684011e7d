"""a19 / f2: the fused optimiser against torch.optim.Adam, configured exactly as the reference does
(scripts/train.py:50-63: Adam, lr 5e-3, betas (0.9, 0.99), eps 1e-15, L2 weight decay 1e-6 on the MLP group only),
and the EMA of torch_ngp/nerf/utils.py:311-315 (torch_ema.ExponentialMovingAverage; the package is a third-party
dependency absent from the reference tree — its published update rule is restated here:
    decay_t = min(decay, (1 + num_updates) / (10 + num_updates));  shadow -= (1 - decay_t) * (shadow - param)).
torch.optim.Adam itself is the oracle (it IS the reference's optimiser), run on the same device in fp32."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _groups(sizes, wd, seed):
    g = torch.Generator().manual_seed(seed)
    ps = [torch.nn.Parameter((torch.randn(n, generator=g) * 0.1).cuda()) for n in sizes]
    return ps, [{'name': 'encoding', 'params': ps[:1]}, {'name': 'net', 'params': ps[1:], 'weight_decay': wd}]


@pytest.mark.parametrize("wd", [0.0, 1e-6, 1e-2])
@pytest.mark.parametrize("sizes", [(4096, 1024), (1000003, 6147, 5), (7, 3)])
def test_fused_adam_matches_torch_adam(sizes, wd):
    """>= 10 steps, sizes that are not multiples of 4 (scalar tail), weight decay on one group only, fresh gradients
    every step, plus a StepLR-style learning-rate change half way."""
    from autolabel_b200.optim import FusedAdam
    pa, ga = _groups(sizes, wd, seed=1)
    pb, gb = _groups(sizes, wd, seed=1)
    ours = FusedAdam(ga, lr=5e-3, betas=(0.9, 0.99), eps=1e-15)
    ref = torch.optim.Adam(gb, lr=5e-3, betas=(0.9, 0.99), eps=1e-15)
    g = torch.Generator().manual_seed(2)
    for step in range(12):
        if step == 6:
            for grp in ours.param_groups + ref.param_groups:
                grp['lr'] *= 0.5
        for a, b in zip(pa, pb):
            # gradients spanning many orders of magnitude (eps = 1e-15 makes tiny gradients matter), some exact zeros
            gr = torch.randn(a.numel(), generator=g) * 10.0 ** torch.randint(-9, 1, (a.numel(),), generator=g).float()
            gr[::13] = 0.0
            a.grad = gr.cuda().clone()
            b.grad = gr.cuda().clone()
        ours.step()
        ref.step()
        for a, b in zip(pa, pb):
            assert torch.all(a.grad == 0), "the fused step zeroes the gradient buffer"
            # a step moves a parameter by ~lr = 5e-3; the two implementations round differently (fused multiply-adds here,
            # torch's lerp / addcdiv there): a few fp32 ulps of the parameter per step, 1e-4 of the motion.  A wrong bias
            # correction, weight decay or epsilon placement shows up at >= 1e-4 absolute.
            assert torch.allclose(a.detach(), b.detach(), rtol=1e-6, atol=5e-7), (step, (a - b).abs().max().item())
    for a, b in zip(pa, pb):
        sa, sb = ours.state[a], ref.state[b]
        # moments are sums of gradients spanning ten decades (cancellation): tolerance relative to the largest entry
        for key in ('exp_avg', 'exp_avg_sq'):
            tol = 2e-6 * sb[key].abs().max().item()
            assert (sa[key] - sb[key]).abs().max().item() <= tol, key


def test_fused_adam_grad_scale_is_unscale():
    """grad_scale (1 / loss scale, 1 / world size) multiplies the gradient before everything else, like
    GradScaler.unscale_ (autolabel/trainer.py:45-48)."""
    from autolabel_b200.optim import FusedAdam
    pa, ga = _groups((515, 64), 1e-6, seed=3)
    pb, gb = _groups((515, 64), 1e-6, seed=3)
    ours = FusedAdam(ga, lr=5e-3, betas=(0.9, 0.99), eps=1e-15)
    ours.grad_scale = 1.0 / 1024.0
    ref = torch.optim.Adam(gb, lr=5e-3, betas=(0.9, 0.99), eps=1e-15)
    g = torch.Generator().manual_seed(4)
    for _ in range(10):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.numel(), generator=g).cuda()
            a.grad = gr * 1024.0
            b.grad = gr.clone()
        ours.step()
        ref.step()
    for a, b in zip(pa, pb):
        assert torch.allclose(a.detach(), b.detach(), rtol=1e-6, atol=5e-7)


def test_multi_tensor_adam_device_step_matches_torch_adam():
    """The graph-resident form (al_adam_multi: one launch for all tensors, step count and learning rate read from device
    memory) against torch.optim.Adam."""
    from autolabel_b200.optim import FusedAdam
    pa, ga = _groups((40004, 6148, 1028), 1e-6, seed=5)
    pb, gb = _groups((40004, 6148, 1028), 1e-6, seed=5)
    ours = FusedAdam(ga, lr=5e-3, betas=(0.9, 0.99), eps=1e-15)
    ref = torch.optim.Adam(gb, lr=5e-3, betas=(0.9, 0.99), eps=1e-15)
    g = torch.Generator().manual_seed(6)
    for step in range(11):
        if step == 5:
            for grp in ours.param_groups + ref.param_groups:
                grp['lr'] *= 0.5
        for a, b in zip(pa, pb):
            gr = torch.randn(a.numel(), generator=g).cuda() * 1e-3
            a.grad = gr.clone() if a.grad is None else a.grad.copy_(gr)
            b.grad = gr.clone()
        ours.step_device()
        ref.step()
    for a, b in zip(pa, pb):
        assert torch.all(a.grad == 0)
        assert torch.allclose(a.detach(), b.detach(), rtol=1e-6, atol=5e-7), (a - b).abs().max().item()
    assert ours.state[pa[0]]['step'] == 11


def test_ema_matches_torch_ema_rule():
    from autolabel_b200.trainer import _EMA
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(1000).cuda()), torch.nn.Parameter(torch.randn(7, 3).cuda())]
    ema = _EMA(ps, 0.95)
    shadow = [p.detach().clone() for p in ps]
    n = 0
    for _ in range(25):
        with torch.no_grad():
            for p in ps:
                p.add_(torch.randn_like(p) * 0.1)
        ema.update()
        n += 1
        d = min(0.95, (1 + n) / (10 + n))
        for s, p in zip(shadow, ps):
            s.sub_((1.0 - d) * (s - p.detach()))
    for s, e in zip(shadow, ema.shadow):
        assert torch.allclose(s, e, rtol=1e-5, atol=1e-6)
    # store / copy_to / restore (torch_ngp/nerf/utils.py:998-1000,1119-1120)
    before = [p.detach().clone() for p in ps]
    ema.store()
    ema.copy_to()
    for s, p in zip(shadow, ps):
        assert torch.allclose(s, p.detach(), rtol=1e-5, atol=1e-6)
    ema.restore()
    for b, p in zip(before, ps):
        assert torch.equal(b, p.detach())
    sd = ema.state_dict()
    assert {'decay', 'num_updates', 'shadow_params'} <= set(sd)      # torch_ema's state_dict keys
