"""Compositing parity: 3-channel subset against the reference's own kernels (oracle/_ref), the
K-channel generalisation (forward + backward incl. the depth gradient) against the fp32 PyTorch
restatement (oracle/field_oracle.py::composite) with autograd.  Tolerance: 1e-3 absolute as stated by
BASELINE.json's north_star (measured errors are ~1e-6)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ATOL = 1e-3  # north_star tolerance for rendered outputs / gradients


def _segments(N, seed, max_len=70, empty_every=7):
    rng = np.random.RandomState(seed)
    counts = rng.randint(1, max_len, size=N)
    counts[::empty_every] = 0
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]])
    perm = rng.permutation(N)  # ray ids are a permutation of the slots
    rays = np.stack([perm, offs, counts], axis=1).astype(np.int32)
    return rays, int(counts.sum())


def _inputs(N, K, seed):
    rays, total = _segments(N, seed)
    M = total + 128 - total % 128
    g = torch.Generator().manual_seed(seed)
    sigmas = (torch.rand(M, generator=g) * 6).cuda()
    vals = torch.randn(M, K, generator=g).cuda()
    deltas = torch.stack([torch.full((M,), 0.0034), torch.rand(M, generator=g) * 0.02 + 0.0034], dim=1).cuda()
    tpos = (torch.rand(M, generator=g) * 5 + 0.2).cuda()
    xyzs = torch.randn(M, 3, generator=g).cuda()
    return torch.from_numpy(rays).cuda(), M, sigmas, vals, deltas, tpos, xyzs


def test_composite_train_3ch_vs_reference_kernels(ref_rm):
    from autolabel_b200 import raymarching as rm
    N, K = 1000, 3
    rays, M, sigmas, rgbs, deltas, _, _ = _inputs(N, K, 5)
    sigmas.requires_grad_(True); rgbs.requires_grad_(True)
    ws, depth, image = rm.composite_rays_train(sigmas, rgbs, deltas, rays)
    rws, rdepth, rimage = torch.empty(N, device='cuda'), torch.empty(N, device='cuda'), torch.empty(N, 3, device='cuda')
    ref_rm.composite_rays_train_forward(sigmas.detach(), rgbs.detach(), deltas, rays, M, N, rws, rdepth, rimage)
    assert torch.allclose(ws, rws, atol=1e-5) and torch.allclose(depth, rdepth, atol=1e-5)
    assert torch.allclose(image, rimage, atol=1e-5)
    g_ws, g_img = torch.randn(N, device='cuda'), torch.randn(N, 3, device='cuda')
    (ws * g_ws).sum().add((image * g_img).sum()).backward()   # no depth term: the reference has none
    rgs, rgr = torch.zeros(M, device='cuda'), torch.zeros(M, 3, device='cuda')
    ref_rm.composite_rays_train_backward(g_ws, g_img, sigmas.detach(), rgbs.detach(), deltas, rays, rws, rimage, M, N, rgs, rgr)
    assert torch.allclose(sigmas.grad, rgs, atol=1e-4, rtol=1e-4)
    assert torch.allclose(rgbs.grad, rgr, atol=1e-5)


@pytest.mark.parametrize("K", [3, 69, 517])
def test_composite_train_k_channels_vs_torch(K):
    from autolabel_b200 import raymarching as rm
    from oracle import field_oracle as fo
    N = 300
    rays, M, sigmas, vals, deltas, tpos, xyzs = _inputs(N, K, 7 + K)
    s1, v1 = sigmas.clone().requires_grad_(True), vals.clone().requires_grad_(True)
    ws, depth, dsq, out, coords = rm.composite_train_full(s1, v1, deltas, rays, tpos=tpos, xyzs=xyzs, sigma_scale=1.3, M=M)
    s2, v2 = sigmas.clone().requires_grad_(True), vals.clone().requires_grad_(True)
    ows, odepth, odsq, oout, ocoords = fo.composite(s2, v2, deltas, tpos, xyzs, rays, M, sigma_scale=1.3)
    for a, b in [(ws, ows), (depth, odepth), (dsq, odsq), (out, oout), (coords, ocoords)]:
        assert (a - b).abs().max().item() < ATOL
        assert (a - b).abs().max().item() < 2e-4   # measured head-room
    g = torch.Generator().manual_seed(1)
    gw, gd, go = torch.randn(N, generator=g).cuda(), torch.randn(N, generator=g).cuda(), torch.randn(N, K, generator=g).cuda()
    ((ws * gw).sum() + (depth * gd).sum() + (out * go).sum()).backward()
    ((ows * gw).sum() + (odepth * gd).sum() + (oout * go).sum()).backward()
    scale = max(1.0, s2.grad.abs().max().item())
    assert (s1.grad - s2.grad).abs().max().item() < ATOL * scale
    assert (v1.grad - v2.grad).abs().max().item() < ATOL


def test_composite_overflow_rays_are_empty():
    from autolabel_b200 import raymarching as rm
    N, K = 64, 5
    rays, M, sigmas, vals, deltas, tpos, xyzs = _inputs(N, K, 11)
    Msmall = M // 2
    ws, depth, dsq, out, coords = rm.composite_train_full(sigmas, vals, deltas, rays, tpos=tpos, xyzs=xyzs, M=Msmall)
    r = rays.cpu().numpy()
    dropped = (r[:, 2] == 0) | (r[:, 1] + r[:, 2] >= Msmall)
    assert dropped.any() and (~dropped).any()
    ids = torch.from_numpy(r[dropped, 0]).long().cuda()
    assert float(ws[ids].abs().sum()) == 0 and float(out[ids].abs().sum()) == 0 and float(depth[ids].abs().sum()) == 0
