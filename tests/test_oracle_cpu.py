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
