"""CPU self-checks of the oracle restatements (no GPU, no golden files)."""
import numpy as np
import torch

from oracle import field_oracle as fo
from oracle import ngp


def test_morton_roundtrip_and_packbits():
    rng = np.random.RandomState(0)
    c = rng.randint(0, 1024, size=(5000, 3)).astype(np.int32)
    m = ngp.morton3D(c)
    assert np.array_equal(ngp.morton3D_invert(m), c)
    assert ngp.morton3D(np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [3, 3, 3]], np.int32)).tolist() == [1, 2, 4, 63]
    g = rng.uniform(0, 0.02, size=4096).astype(np.float32)
    b = ngp.packbits(g, 0.01)
    assert np.array_equal(np.unpackbits(b, bitorder='little').astype(bool), g > 0.01)


def test_march_full_occupancy_uniform_steps():
    """Fully occupied grid: every ray gets min(max_steps, ceil((far-near)/dt)) samples spaced dt = 2 sqrt(3)/1024."""
    N = 64
    rng = np.random.RandomState(1)
    o = rng.uniform(-0.5, 0.5, size=(N, 3)).astype(np.float32)
    d = rng.normal(size=(N, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    bits = np.full(128 ** 3 // 8, 255, np.uint8)
    n, f, _, _ = ngp.near_far_from_aabb(o, d, np.array([-1, -1, -1, 1, 1, 1], np.float32), 0.2)
    r = ngp.march_rays_train(o, d, 1.0, bits, 1, 128, n, f, M=N * 1024, perturb=False)
    dt = np.float32(2 * np.float32(1.7320508075688772) / 1024)
    cnt = r["rays"][:, 2]
    assert np.all(np.abs(cnt - np.minimum(1024, np.ceil((f - n) / dt))) <= 1)
    assert np.all(r["deltas"][: cnt[0], 0] == dt)
    assert np.array_equal(r["rays"][:, 1], np.concatenate([[0], np.cumsum(cnt)[:-1]]))
    # empty grid: no samples
    r0 = ngp.march_rays_train(o, d, 1.0, np.zeros_like(bits), 1, 128, n, f, M=1024, perturb=True)
    assert r0["counter"][0] == 0 and r0["counter"][1] == N


def test_composite_backward_matches_autograd():
    rng = np.random.RandomState(2)
    N, K = 20, 5
    counts = rng.randint(0, 30, size=N)
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]])
    rays = np.stack([rng.permutation(N), offs, counts], 1).astype(np.int32)
    M = int(counts.sum()) + 5
    sig = rng.uniform(0, 5, M).astype(np.float32); vals = rng.normal(size=(M, K)).astype(np.float32)
    dl = np.stack([np.full(M, 0.01), rng.uniform(0.01, 0.02, M)], 1).astype(np.float32)
    ws, depth, img = ngp.composite_rays_train_forward(sig, vals, dl, rays, M)
    s, v = torch.tensor(sig, requires_grad=True), torch.tensor(vals, requires_grad=True)
    tpos = torch.zeros(M)
    ows, _, _, oout, _ = fo.composite(s, v, torch.tensor(dl), tpos, torch.zeros(M, 3), torch.tensor(rays), M)
    assert np.abs(ows.detach().numpy() - ws).max() < 1e-5 and np.abs(oout.detach().numpy() - img).max() < 1e-5
    gws, gim = rng.normal(size=N).astype(np.float32), rng.normal(size=(N, K)).astype(np.float32)
    ((ows * torch.tensor(gws)).sum() + (oout * torch.tensor(gim)).sum()).backward()
    gs, gv = ngp.composite_rays_train_backward(gws, gim, sig, vals, dl, rays, ws, img, M)
    assert np.abs(gs - s.grad.numpy()).max() < 1e-4 and np.abs(gv - v.grad.numpy()).max() < 1e-5


def test_torch_grid_matches_c_grid():
    offsets = ngp.grid_offsets(16, 16, 2.0, 19, 3)
    rng = np.random.RandomState(3)
    x = rng.uniform(0, 1, size=(300, 3)).astype(np.float32); x[:5] = -0.1
    table = rng.uniform(-1, 1, size=(int(offsets[-1]), 2)).astype(np.float32)
    out, _ = ngp.grid_encode_forward(x, table, offsets, 2.0, 16, 0)
    t = fo.grid_encode(torch.tensor(x), torch.tensor(table), offsets, 2.0, 16, 0).numpy()
    assert np.abs(t - out.transpose(1, 0, 2).reshape(300, 32)).max() < 1e-5


def test_field_oracle_shapes_and_padding():
    torch.manual_seed(0)
    n = 50
    x = torch.randn(n, 44)
    p = torch.randn(fo.mlp_num_params(48, 128, 16, 2)) * 0.1
    y = fo.mlp(x, p, 48, 128, 16, 2)
    assert y.shape == (n, 16)
    # the padded columns are ones: changing the weights of a padded column shifts the output like a bias
    p2 = p.clone(); p2.view(-1)[: 128 * 48].view(128, 48)[:, 47] += 1.0
    assert not torch.allclose(fo.mlp(x, p2, 48, 128, 16, 2), y)
    e = fo.freq_encode(torch.tensor([[0.25, 0.5, 1.0]]), 2)
    assert e.shape == (1, 12)
    assert torch.allclose(e[0, :4], torch.tensor([np.sin(np.pi / 4), np.cos(np.pi / 4), 1.0, 0.0], dtype=torch.float32), atol=1e-6)
    sh = fo.sh4(torch.tensor([[0.5, 0.5, 1.0]]))   # direction (0,0,1)
    assert abs(sh[0, 0].item() - 0.28209479) < 1e-6 and abs(sh[0, 2].item() - 0.48860251) < 1e-6


def test_pose_gradients_of_slab_test_and_march():
    """The reference fork's backward passes w.r.t. the rays (raymarching.py:81-136, :358-392), pure torch in
    autolabel_b200.raymarching: checked against autograd through a torch restatement of the forward formulas."""
    from autolabel_b200 import raymarching as rm      # the two backward helpers are pure torch (no kernel involved)
    g = torch.Generator().manual_seed(0)
    N, bound = 200, 2.0
    o = (torch.rand(N, 3, generator=g) - 0.5) * 1.5 * bound
    d = torch.nn.functional.normalize(torch.randn(N, 3, generator=g), dim=1)
    d[:20] *= 0.0
    d[:20, 0] = 1.0                      # axis-parallel rays: zero components
    d[:20] += 1e-3                       # (kept finite: the reference divides by d)
    aabb = torch.tensor([-bound] * 3 + [bound] * 3)
    nears, fars, ni, fi = ngp.near_far_from_aabb(o.numpy(), d.numpy(), aabb.numpy(), 0.2)
    ni, fi = torch.from_numpy(ni.astype(np.int64)), torch.from_numpy(fi.astype(np.int64))
    o1, d1 = o.clone().requires_grad_(True), d.clone().requires_grad_(True)
    rows = torch.arange(N)
    hit = (ni != 255)

    def plane_t(idx):
        ax = idx % 3
        return (aabb[idx % 6] - o1[rows, ax]) / d1[rows, ax]
    gn, gf = torch.randn(N, generator=g), torch.randn(N, generator=g)
    loss = ((plane_t(ni) * gn + plane_t(fi) * gf) * hit).sum()
    loss.backward()
    g_o, g_d = rm.near_far_backward(aabb, o, d, ni, fi, gn, gf)
    assert torch.allclose(g_o, o1.grad, rtol=1e-5, atol=1e-6) and torch.allclose(g_d, d1.grad, rtol=1e-5, atol=1e-5)
    assert hit.sum() > 50 and (~hit).sum() >= 0

    # march: ragged segments in ray order with padding and one dropped (count 0) ray
    counts = torch.randint(0, 9, (N,), generator=g)
    counts[5] = 0
    offsets = torch.cumsum(counts, 0) - counts
    total = int(counts.sum())
    M = total + 17
    rays = torch.stack([torch.arange(N), offsets, counts], 1).int()
    ts = torch.rand(M, 1, generator=g) * 3
    ray_of = torch.repeat_interleave(torch.arange(N), counts)
    o2, d2 = o.clone().requires_grad_(True), d.clone().requires_grad_(True)
    xyz = o2[ray_of] + ts[:total] * d2[ray_of]
    dirs = d2[ray_of]
    G1, G2 = torch.randn(M, 3, generator=g), torch.randn(M, 3, generator=g)
    ((xyz * G1[:total]).sum() + (dirs * G2[:total]).sum()).backward()
    g_o, g_d = rm.march_backward(rays, ts, G1, G2, N)
    assert torch.allclose(g_o, o2.grad, rtol=1e-5, atol=1e-5) and torch.allclose(g_d, d2.grad, rtol=1e-5, atol=1e-5)
