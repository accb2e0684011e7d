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


def _inputs(N, K, seed, max_len=70):
    rays, total = _segments(N, seed, max_len=max_len)
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


def test_composite_wide_rows_long_rays():
    """The C5 shape of compositing (517 value channels, few rays, several hundred samples per ray, some beyond 1024):
    the forward runs as channel slices, the rank-1 backward as a chunk-parallel dot-product pass + the per-ray scan.
    Against the fp32 torch restatement with autograd."""
    from autolabel_b200 import raymarching as rm
    from oracle import field_oracle as fo
    N, K = 48, 517
    rays, M, sigmas, vals, deltas, tpos, xyzs = _inputs(N, K, 31, max_len=1500)
    sigmas = sigmas * 0.02                       # long rays: keep the transmittance alive along the whole ray
    s1, v1 = sigmas.clone().requires_grad_(True), vals.clone().requires_grad_(True)
    ws, depth, dsq, out, coords = rm.composite_train_full(s1, v1, deltas, rays, tpos=tpos, xyzs=xyzs, sigma_scale=1.3, M=M)
    s2, v2 = sigmas.clone().requires_grad_(True), vals.clone().requires_grad_(True)
    ows, odepth, odsq, oout, ocoords = fo.composite(s2, v2, deltas, tpos, xyzs, rays, M, sigma_scale=1.3)
    for a, b in [(ws, ows), (depth, odepth), (dsq, odsq), (out, oout), (coords, ocoords)]:
        assert (a - b).abs().max().item() < 2e-4 * max(1.0, b.abs().max().item())
    g = torch.Generator().manual_seed(3)
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


@pytest.mark.parametrize("K,N,max_len", [(3, 300, 70), (69, 300, 70), (517, 300, 70), (517, 48, 1500)])
def test_rank1_backward_equals_materialised(K, N, max_len):
    """al_composite_train_bwd_weights (w, dL/dsigma per sample) reproduces al_composite_train_bwd:
    dL/dvals[i, c] == w[i] * g_out[ray(i), c].  The last case is the C5 shape (wide rows, few long rays, some beyond
    1024 samples), which the weights form runs as a chunk-parallel dot-product pass followed by the per-ray scan."""
    from autolabel_b200._lib import call, ptr, stream_ptr
    rays, M, sigmas, vals, deltas, tpos, xyzs = _inputs(N, K, 21 + K, max_len=max_len)
    if max_len > 100:
        sigmas = sigmas * 0.02
    dev = sigmas.device
    st = stream_ptr(dev)
    f32 = dict(dtype=torch.float32, device=dev)
    ws, depth, dsq = torch.empty(N, **f32), torch.empty(N, **f32), torch.empty(N, **f32)
    out, coords = torch.empty(N, K, **f32), torch.empty(N, 3, **f32)
    call("al_composite_train_fwd", ptr(sigmas), 1, ptr(vals), K, K, ptr(deltas), ptr(tpos), ptr(xyzs), ptr(rays), M, N,
         1.3, ptr(ws), ptr(depth), ptr(dsq), ptr(out), ptr(coords), st)
    g = torch.Generator().manual_seed(2)
    gw, gd, go = torch.randn(N, generator=g).cuda(), torch.randn(N, generator=g).cuda(), torch.randn(N, K, generator=g).cuda()
    gs_a, gv_a, am_a = torch.zeros(M, **f32), torch.zeros(M, K, **f32), torch.zeros(1, **f32)
    call("al_composite_train_bwd", ptr(gw), ptr(gd), ptr(go), ptr(sigmas), 1, ptr(vals), K, K, ptr(deltas), ptr(tpos),
         ptr(rays), ptr(ws), ptr(depth), ptr(out), M, N, 1.3, ptr(gs_a), 1, ptr(gv_a), K, ptr(am_a), st)
    w_b, gs_b, am_b = torch.zeros(M, **f32), torch.zeros(M, **f32), torch.zeros(1, **f32)
    call("al_composite_train_bwd_weights", ptr(gw), ptr(gd), ptr(go), ptr(sigmas), 1, ptr(vals), K, K, ptr(deltas),
         ptr(tpos), ptr(rays), ptr(ws), ptr(depth), ptr(out), M, N, 1.3, ptr(w_b), ptr(gs_b), ptr(am_b), st)
    r = rays.cpu().numpy()
    sray = torch.zeros(M, dtype=torch.long)
    for rid, off, cnt in r:
        sray[off:off + cnt] = rid
    gv_b = w_b[:, None] * go[sray.cuda()]
    live = int(r[:, 2].sum())
    scale = max(1.0, gs_a.abs().max().item())
    assert (gs_a[:live] - gs_b[:live]).abs().max().item() < 1e-4 * scale
    assert (gv_a[:live] - gv_b[:live]).abs().max().item() < 1e-5
    assert am_b.item() >= 0.999 * am_a.item() and am_b.item() < 8 * am_a.item()
