"""CPU port of the path the reference executes today: ``NeRFRenderer.run()`` (uniform sampling +
PyTorch compositing, torch_ngp/nerf/renderer.py:186-320) driven by ``SimpleTrainer.train_step``
(autolabel/trainer.py:54-94) with the optimiser of scripts/train.py:50-63, in fp32 PyTorch with the
field of oracle/field_oracle.py and the slab test of oracle/ngp_oracle.c.

PINNED on the reference's own code: tests/golden/ref_run_path.npz holds outputs of the UNMODIFIED
autolabel.models.ALNetwork.run and autolabel.trainer.SimpleTrainer.train_step (imported from the reference tree
on CPU by tests/golden/make_golden_run.py, tiny-cuda-nn replaced by a shim over field_oracle.py); run() and
loss_fn() below reproduce the six output maps and the loss to 2e-5 (tests/test_oracle_pinned.py).

TEST INFRASTRUCTURE ONLY.  Used as (a) tier O3 reference of the run() path and (b) the CPU baseline that
bench.py reports (`cpu_baseline`, kind "port") and times under `--impl reference`: the reference's
own Python cannot travel to the GPU box (/root/reference is absent there) and imports tiny-cuda-nn,
which does not exist in this image, so this port stands in for it.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import field_oracle as fo
from . import ngp


class OracleField:
    """fp32 parameters of an ALNetwork (same flat layouts as autolabel_b200.models.ALNetwork)."""

    def __init__(self, encoding='hg+freq', hidden=128, hidden_color=128, feat_dim=64, n_classes=2, bound=1.0,
                 density_scale=1.0, min_near=0.2, seed=0, device='cpu'):
        g = torch.Generator().manual_seed(seed)
        self.cfg = dict(encoding=encoding, bound=float(bound), hidden=hidden, hidden_color=hidden_color,
                        feat_dim=feat_dim, n_classes=n_classes, per_level_scale=2.0, H=16)
        self.density_scale, self.min_near, self.bound = density_scale, min_near, float(bound)
        width = {'freq': 60, 'hg': 32, 'hg+freq': 44}[encoding]
        in_pad = fo.pad16(width)

        def xavier(shapes):
            return torch.cat([(torch.rand(o * i, generator=g) * 2 - 1) * math.sqrt(6.0 / (o + i)) for o, i in shapes])

        P = {
            'w_sigma': xavier([(hidden, in_pad), (hidden, hidden), (16, hidden)]),
            'w_color': xavier([(hidden_color, 32), (hidden_color, hidden_color), (16, hidden_color)]),
            'w_semf': xavier([(feat_dim, 16), (feat_dim, feat_dim), (feat_dim, feat_dim)]),
            'w_semo': xavier([(64, feat_dim + 16), (16, 64)]),
        }
        if encoding != 'freq':
            pls = 2.0 if encoding == 'hg+freq' else float(np.exp2(np.log2(2 ** 18 / 16) / 15))
            self.cfg['per_level_scale'] = pls
            self.cfg['offsets'] = fo.grid_offsets(16, 16, pls, 19, 3)
            P['table'] = (torch.rand(int(self.cfg['offsets'][-1]), 2, generator=g) * 2 - 1) * 1e-4
        self.P = {k: v.to(device).requires_grad_(True) for k, v in P.items()}
        self.device = device

    def parameters(self):
        return list(self.P.values())

    def optimizer(self, lr=5e-3):
        groups = []
        if 'table' in self.P:
            groups.append({'params': [self.P['table']]})
        groups.append({'params': [self.P[k] for k in ('w_sigma', 'w_color', 'w_semf', 'w_semo')], 'weight_decay': 1e-6})
        return torch.optim.Adam(groups, lr=lr, betas=(0.9, 0.99), eps=1e-15)


def run(field, rays_o, rays_d, direction_norms, num_steps=256, perturb=False):
    """renderer.py:186-320 (uniform z samples, weights via cumprod with the 1e-15 guard, 1e-4 weight mask for
    the colour query, white background, semantic heads on all samples)."""
    N = rays_o.shape[0]
    dev = rays_o.device
    b = field.bound
    aabb = np.array([-b, -b, -b, b, b, b], np.float32)
    nears, fars, _, _ = ngp.near_far_from_aabb(rays_o.detach().cpu().numpy(), rays_d.detach().cpu().numpy(), aabb, field.min_near)
    nears = torch.from_numpy(nears).to(dev).unsqueeze(-1)
    fars = torch.from_numpy(fars).to(dev).unsqueeze(-1)
    z = torch.linspace(0.0, 1.0, num_steps, device=dev).unsqueeze(0).expand(N, num_steps)
    z = nears + (fars - nears) * z
    sample_dist = (fars - nears) / num_steps
    if perturb:
        z = z + (torch.rand(z.shape, device=dev) - 0.5) * sample_dist
    xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * z.unsqueeze(-1)
    xyzs = torch.min(torch.max(xyzs, torch.full((3,), -b, device=dev)), torch.full((3,), b, device=dev))
    flat = xyzs.reshape(-1, 3)
    dirs = rays_d.view(-1, 1, 3).expand_as(xyzs).reshape(-1, 3)
    sigma, rgb, logits, feat, _ = fo.field_forward(flat, dirs, field.P, field.cfg)
    sigma = sigma.view(N, num_steps)
    deltas = torch.cat([z[..., 1:] - z[..., :-1], sample_dist * torch.ones_like(z[..., :1])], dim=-1)
    alphas = 1 - torch.exp(-deltas * field.density_scale * sigma)
    shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-15], dim=-1)
    weights = alphas * torch.cumprod(shifted, dim=-1)[..., :-1]
    mask = weights > 1e-4
    rgbs = rgb.view(N, num_steps, 3) * mask.unsqueeze(-1)        # colour is only evaluated where mask (zeros elsewhere)
    weights = weights * mask
    ws = weights.sum(-1)
    norms = direction_norms.reshape(-1)
    depth = (weights * z).sum(-1) / norms
    depth_variance = (weights * (depth[..., None] - z) ** 2).sum(-1).detach()
    w = weights.unsqueeze(-1)
    C, Fd = field.cfg['n_classes'], field.cfg['feat_dim']
    return {
        'depth': depth, 'depth_variance': depth_variance,
        'image': (w * rgbs).sum(-2) + (1 - ws).unsqueeze(-1),
        'semantic': (w * logits.view(N, num_steps, C)).sum(-2),
        'semantic_features': (w * feat.view(N, num_steps, Fd)).sum(-2),
        'coordinates_map': (w * xyzs).sum(-2),
    }


def loss_fn(outputs, data, rgb_weight=1.0, depth_weight=0.1, feature_weight=0.5, semantic_weight=1.0):
    """autolabel/trainer.py:72-92."""
    loss = rgb_weight * ((outputs['image'] - data['pixels']) ** 2).mean()
    has_depth = data['depth'] > 0.01
    if has_depth.any():
        loss = loss + depth_weight * torch.abs(outputs['depth'][has_depth] - data['depth'][has_depth]).mean()
    if 'features' in data:
        gt = data['features']
        loss = loss + feature_weight * F.l1_loss(outputs['semantic_features'][:, :gt.shape[1]], gt)
    has_sem = data['semantic'] >= 0
    if has_sem.any():
        loss = loss + semantic_weight * F.cross_entropy(outputs['semantic'][has_sem], data['semantic'][has_sem])
    return loss


def train_step(field, optimizer, data, num_steps=256):
    """One reference-style iteration (trainer.py:40-48 without AMP): zero_grad, render, loss, backward, step."""
    optimizer.zero_grad()
    out = run(field, data['rays_o'], data['rays_d'], data['direction_norms'], num_steps=num_steps, perturb=True)
    loss = loss_fn(out, data)
    loss.backward()
    optimizer.step()
    return float(loss.item())
