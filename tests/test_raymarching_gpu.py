"""Tier O1 (bit-exact) parity of the geometry / marching kernels against the reference's own
kernels (oracle/_ref, compiled unmodified from torch_ngp/raymarching/src/raymarching.cu) and the
CPU restatement (oracle/ngp_oracle.c), all called on identical seeded inputs."""
import numpy as np
import pytest
import torch

from tests.helpers import aabb_of, make_density_grid, make_rays

pytestmark = pytest.mark.gpu

BOUND = 3.0
CASCADE = 3
H = 128


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def scene():
    from autolabel_b200 import raymarching as rm
    o, d = make_rays(4096, BOUND, seed=1, inside=False)
    grid = make_density_grid(CASCADE, H, seed=2)
    bits = rm.packbits(_dev(grid), 0.01)
    return dict(o=o, d=d, grid=grid, bits=bits)


def test_near_far_bit_exact(ref_rm, scene):
    from autolabel_b200 import raymarching as rm
    from oracle import ngp
    o, d = _dev(scene['o']), _dev(scene['d'])
    aabb = _dev(aabb_of(BOUND))
    n, f, ni, fi = rm.near_far_from_aabb(o, d, aabb, 0.2, return_indices=True)
    N = o.shape[0]
    rn, rf = torch.empty(N, device='cuda'), torch.empty(N, device='cuda')
    rni, rfi = torch.empty(N, dtype=torch.uint8, device='cuda'), torch.empty(N, dtype=torch.uint8, device='cuda')
    ref_rm.near_far_from_aabb(o, d, aabb, N, 0.2, rn, rf, rni, rfi)
    torch.cuda.synchronize()
    assert torch.equal(n.view(torch.int32), rn.view(torch.int32))
    assert torch.equal(f.view(torch.int32), rf.view(torch.int32))
    assert torch.equal(ni, rni) and torch.equal(fi, rfi)
    assert (ni == 255).any(), "the test set must contain rays that miss the box"
    cn, cf, cni, cfi = ngp.near_far_from_aabb(scene['o'], scene['d'], aabb_of(BOUND), 0.2)
    assert np.array_equal(cn.view(np.int32), n.cpu().numpy().view(np.int32))
    assert np.array_equal(cf.view(np.int32), f.cpu().numpy().view(np.int32))
    assert np.array_equal(cni, ni.cpu().numpy()) and np.array_equal(cfi, fi.cpu().numpy())


def test_morton_bit_exact(ref_rm):
    from autolabel_b200 import raymarching as rm
    from oracle import ngp
    rng = np.random.RandomState(3)
    coords = rng.randint(0, 128, size=(100003, 3)).astype(np.int32)
    c = _dev(coords)
    ind = rm.morton3D(c)
    rind = torch.empty_like(ind)
    ref_rm.morton3D(c, c.shape[0], rind)
    assert torch.equal(ind, rind)
    back = rm.morton3D_invert(ind)
    rback = torch.empty_like(back)
    ref_rm.morton3D_invert(ind, ind.shape[0], rback)
    assert torch.equal(back, rback) and torch.equal(back, c)
    assert np.array_equal(ngp.morton3D(coords), ind.cpu().numpy())
    assert np.array_equal(ngp.morton3D_invert(ind.cpu().numpy()), coords)


@pytest.mark.parametrize("thresh", [0.01, 0.0, 0.5])
def test_packbits_bit_exact(ref_rm, scene, thresh):
    from autolabel_b200 import raymarching as rm
    from oracle import ngp
    g = _dev(scene['grid'])
    bits = rm.packbits(g, thresh)
    rbits = torch.empty_like(bits)
    ref_rm.packbits(g, bits.numel(), thresh, rbits)
    assert torch.equal(bits, rbits)
    assert np.array_equal(ngp.packbits(scene['grid'], thresh), bits.cpu().numpy())
    # device-side threshold: min(thresh, *thresh_dev)
    td = torch.tensor([0.004], device='cuda')
    b2 = rm.packbits(g, thresh, thresh_dev=td)
    assert np.array_equal(ngp.packbits(scene['grid'], min(thresh, 0.004)), b2.cpu().numpy())


def _ref_march_train(ref_rm, o, d, bits, nears, fars, M, perturb, dt_gamma=0.0, max_steps=1024):
    N = o.shape[0]
    xyzs = torch.zeros(M, 3, device='cuda'); dirs = torch.zeros(M, 3, device='cuda')
    deltas = torch.zeros(M, 2, device='cuda'); ts = torch.zeros(M, 1, device='cuda')
    rays = torch.empty(N, 3, dtype=torch.int32, device='cuda')
    counter = torch.zeros(2, dtype=torch.int32, device='cuda')
    ref_rm.march_rays_train(o, d, bits, BOUND, dt_gamma, max_steps, N, CASCADE, H, M, nears, fars, xyzs, dirs,
                            deltas, ts, rays, counter, 1 if perturb else 0)
    torch.cuda.synchronize()
    return xyzs, dirs, deltas, ts, rays, counter


def _per_ray(rays, arrs, M):
    """{ray id: tuple of that ray's rows} for rays that were written."""
    rays = rays.cpu().numpy()
    arrs = [a.cpu().numpy() for a in arrs]
    out = {}
    for rid, off, cnt in rays:
        if cnt == 0 or off + cnt >= M:
            out[int(rid)] = (int(cnt), None)
        else:
            out[int(rid)] = (int(cnt), tuple(a[off:off + cnt].copy() for a in arrs))
    return out


@pytest.mark.parametrize("perturb,dt_gamma", [(True, 0.0), (False, 0.0), (True, 1.0 / 256)])
def test_march_rays_train_bit_exact(ref_rm, scene, perturb, dt_gamma):
    """Per-ray sample counts and positions, dirs, deltas, ts: bit-exact per ray id (slot order of the
    reference depends on atomic scheduling, raymarching.cu:448-449)."""
    from autolabel_b200 import raymarching as rm
    from autolabel_b200.raymarching import _march_train_raw
    from oracle import ngp
    o, d = _dev(scene['o']), _dev(scene['d'])
    aabb = _dev(aabb_of(BOUND))
    nears, fars = rm.near_far_from_aabb(o, d, aabb, 0.2)
    N = o.shape[0]
    M = N * 1024
    r = _march_train_raw(o, d, BOUND, scene['bits'], CASCADE, H, nears, fars, None, M, perturb, dt_gamma, 1024,
                         want_tpos=True, want_sray=True)
    rx, rd, rdl, rts, rrays, rcounter = _ref_march_train(ref_rm, o, d, scene['bits'], nears, fars, M, perturb, dt_gamma)
    torch.cuda.synchronize()
    assert int(r['counter'][0]) == int(rcounter[0]) and int(r['counter'][1]) == int(rcounter[1]) == N
    assert int(rcounter[0]) > 10000, "scene too empty for a meaningful test"
    mine = _per_ray(r['rays'], [r['xyzs'], r['dirs'], r['deltas'], r['ts']], M)
    ref = _per_ray(rrays, [rx, rd, rdl, rts], M)
    assert mine.keys() == ref.keys()
    for rid in ref:
        assert mine[rid][0] == ref[rid][0], f"ray {rid}: count {mine[rid][0]} != {ref[rid][0]}"
        if ref[rid][1] is not None:
            for a, b in zip(mine[rid][1], ref[rid][1]):
                assert np.array_equal(a.view(np.int32), b.view(np.int32)), f"ray {rid} differs"
    # deterministic layout of this implementation: offsets = exclusive scan in ray order
    rays = r['rays'].cpu().numpy()
    assert np.array_equal(rays[:, 0], np.arange(N))
    assert np.array_equal(rays[:, 1], np.concatenate([[0], np.cumsum(rays[:, 2])[:-1]]))
    # extras: tpos + dt == ts, sray == ray id
    tot = int(rays[:, 2].sum())
    tp, ts, dl = r['tpos'][:tot], r['ts'][:tot, 0], r['deltas'][:tot, 0]
    assert torch.equal((tp + dl).view(torch.int32), ts.view(torch.int32))
    assert np.array_equal(r['sray'][:tot].cpu().numpy(), np.repeat(np.arange(N), rays[:, 2]))
    # CPU restatement == GPU, slot for slot (both are sequential-scan layouts)
    c = ngp.march_rays_train(scene['o'], scene['d'], BOUND, scene['bits'].cpu().numpy(), CASCADE, H,
                             nears.cpu().numpy(), fars.cpu().numpy(), M=tot + 1, perturb=perturb,
                             dt_gamma=dt_gamma)
    assert np.array_equal(c['rays'], rays)
    assert np.array_equal(c['xyzs'][:tot].view(np.int32), r['xyzs'][:tot].cpu().numpy().view(np.int32))
    assert np.array_equal(c['deltas'][:tot].view(np.int32), r['deltas'][:tot].cpu().numpy().view(np.int32))
    assert np.array_equal(c['ts'][:tot].view(np.int32), r['ts'][:tot, 0].cpu().numpy().view(np.int32))


def test_march_rays_train_overflow_and_fused_slab(ref_rm, scene):
    """Sample budget: rays whose segment does not fit (offset + count >= M) are dropped, counts stay exact;
    and the fused slab test (nears/fars = NULL) gives the same samples."""
    from autolabel_b200 import raymarching as rm
    from autolabel_b200.raymarching import _march_train_raw
    o, d = _dev(scene['o']), _dev(scene['d'])
    aabb = _dev(aabb_of(BOUND))
    nears, fars = rm.near_far_from_aabb(o, d, aabb, 0.2)
    N = o.shape[0]
    full = _march_train_raw(o, d, BOUND, scene['bits'], CASCADE, H, nears, fars, None, N * 1024, True, 0.0, 1024)
    total = int(full['counter'][0])
    M = (total // 2) // 128 * 128
    part = _march_train_raw(o, d, BOUND, scene['bits'], CASCADE, H, nears, fars, None, M, True, 0.0, 1024)
    assert torch.equal(part['rays'], full['rays'])
    assert int(part['meta'][1]) == total
    nv = int(part['meta'][0])
    rays = full['rays'].cpu().numpy()
    ok = (rays[:, 2] > 0) & (rays[:, 1] + rays[:, 2] < M)
    first_bad = np.nonzero((rays[:, 2] > 0) & ~ok)[0][0]
    assert nv == rays[first_bad, 1]
    assert torch.equal(part['xyzs'][:nv], full['xyzs'][:nv])
    assert float(part['xyzs'][nv:].abs().sum()) == 0.0  # nothing written past the budget
    fused = _march_train_raw(o, d, BOUND, scene['bits'], CASCADE, H, None, None, None, N * 1024, True, 0.0, 1024,
                             aabb=aabb, min_near=0.2)
    assert torch.equal(fused['rays'], full['rays'])
    assert torch.equal(fused['xyzs'][:total], full['xyzs'][:total])
    assert torch.equal(fused['nears'].view(torch.int32), nears.view(torch.int32))


def test_march_rays_train_device_budget(scene):
    """al_march_rays_train_budget: capacity M with the budget in device memory == al_march_rays_train with M = budget
    (same live prefix, same samples, same counters); dropped rays are reported with count 0; the budget can change
    between launches without touching the launch arguments."""
    from autolabel_b200 import _lib
    from autolabel_b200 import raymarching as rm
    from autolabel_b200._lib import call, ptr, stream_ptr
    from autolabel_b200.raymarching import _march_train_raw
    o, d = _dev(scene['o']), _dev(scene['d'])
    nears, fars = rm.near_far_from_aabb(o, d, _dev(aabb_of(BOUND)), 0.2)
    N = o.shape[0]
    cap = N * 1024
    full = _march_train_raw(o, d, BOUND, scene['bits'], CASCADE, H, nears, fars, None, cap, True, 0.0, 1024)
    total = int(full['counter'][0])
    dev = o.device
    ws = torch.empty(_lib.lib.al_march_rays_train_workspace(N, 1024), dtype=torch.uint8, device=dev)
    budget = torch.zeros(1, dtype=torch.int32, device=dev)
    xyzs, deltas = torch.zeros(cap, 3, device=dev), torch.zeros(cap, 2, device=dev)
    rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
    counter, meta = torch.zeros(2, dtype=torch.int32, device=dev), torch.zeros(2, dtype=torch.int32, device=dev)
    for frac in (0.5, 0.25, 2.0):
        M = max(128, int(total * frac) // 128 * 128)
        ref = _march_train_raw(o, d, BOUND, scene['bits'], CASCADE, H, nears, fars, None, min(M, cap), True, 0.0, 1024)
        budget.fill_(M)
        counter.zero_(); xyzs.zero_(); deltas.zero_()
        call("al_march_rays_train_budget", ptr(o), ptr(d), ptr(scene['bits']), BOUND, 0.0, 1024, N, CASCADE, H, cap,
             ptr(budget), ptr(nears), ptr(fars), None, 0.2, None, None, ptr(xyzs), None, ptr(deltas), None, None, None,
             ptr(rays), ptr(counter), ptr(meta), 1, ptr(ws), stream_ptr(dev))
        nv = int(ref['meta'][0])
        assert int(meta[0]) == nv and int(meta[1]) == total and int(counter[0]) == total
        assert torch.equal(xyzs[:nv], ref['xyzs'][:nv]) and torch.equal(deltas[:nv], ref['deltas'][:nv])
        assert float(xyzs[nv:].abs().sum()) == 0.0
        r_ref, r_b = ref['rays'].cpu().numpy(), rays.cpu().numpy()
        kept = (r_ref[:, 2] > 0) & (r_ref[:, 1].astype(np.int64) + r_ref[:, 2] < min(M, cap))
        assert np.array_equal(r_b[:, :2], r_ref[:, :2])
        assert np.array_equal(r_b[kept, 2], r_ref[kept, 2]) and (r_b[~kept, 2] == 0).all()


def test_march_rays_train_api(scene):
    """Reference-shaped wrapper: sizing by mean_count / align, truncation to the counted total."""
    from autolabel_b200 import raymarching as rm
    o, d = _dev(scene['o']), _dev(scene['d'])
    nears, fars = rm.near_far_from_aabb(o, d, _dev(aabb_of(BOUND)), 0.2)
    counter = torch.zeros(2, dtype=torch.int32, device='cuda')
    xyzs, dirs, deltas, rays = rm.march_rays_train(o, d, BOUND, scene['bits'], CASCADE, H, nears, fars, counter, -1,
                                                   True, 128, False, 0, 1024)
    total = int(counter[0])
    assert xyzs.shape[0] == total + 128 - total % 128 and dirs.shape == xyzs.shape and deltas.shape[1] == 2
    counter.zero_()
    xyzs2, _, _, _ = rm.march_rays_train(o, d, BOUND, scene['bits'], CASCADE, H, nears, fars, counter, total, True,
                                         128, False, 0, 1024)
    assert xyzs2.shape[0] == total + 128 - total % 128


def test_inference_loop_bit_exact(ref_rm, scene):
    """march_rays / composite_rays / compact_rays against the reference kernels over a full
    reference-style inference loop (renderer.py:403-472) with a synthetic field."""
    from autolabel_b200 import raymarching as rm
    from oracle import ngp
    o, d = _dev(scene['o'][:2048]), _dev(scene['d'][:2048])
    aabb = _dev(aabb_of(BOUND))
    nears, fars = rm.near_far_from_aabb(o, d, aabb, 0.2)
    N = o.shape[0]

    def field(xyzs):
        s = (torch.sin(xyzs * 3.0).sum(-1) + 1.5).clamp(min=0) * 4.0
        rgb = torch.sigmoid(xyzs)
        return s.contiguous(), rgb.contiguous()

    def loop(march, composite, compact):
        ws = torch.zeros(N, device='cuda'); depth = torch.zeros(N, device='cuda'); image = torch.zeros(N, 3, device='cuda')
        n_alive = N
        alive = torch.zeros(2, N, dtype=torch.int32, device='cuda'); rt = torch.zeros(2, N, device='cuda')
        counter = torch.zeros(1, dtype=torch.int32, device='cuda')
        step, i, hist = 0, 0, []
        while step < 1024:
            if step == 0:
                alive[0] = torch.arange(N, dtype=torch.int32, device='cuda'); rt[0] = nears
            else:
                counter.zero_()
                compact(n_alive, alive[i % 2], alive[(i + 1) % 2], rt[i % 2], rt[(i + 1) % 2], counter)
                n_alive = int(counter.item())
            if n_alive <= 0:
                break
            n_step = max(min(N // n_alive, 8), 1)
            xyzs, dirs, deltas = march(n_alive, n_step, alive[i % 2], rt[i % 2])
            s, rgb = field(xyzs)
            composite(n_alive, n_step, alive[i % 2], rt[i % 2], s, rgb, deltas, ws, depth, image)
            hist.append((n_alive, n_step, xyzs.clone(), deltas.clone()))
            step += n_step
            i += 1
        return ws, depth, image, hist

    def my_march(n_alive, n_step, al, t):
        return rm.march_rays(n_alive, n_step, al, t, o, d, BOUND, scene['bits'], CASCADE, H, nears, fars, 128, False, 0, 1024)

    def ref_march(n_alive, n_step, al, t):
        M = n_alive * n_step
        M += 128 - M % 128
        x = torch.zeros(M, 3, device='cuda'); dd = torch.zeros(M, 3, device='cuda'); dl = torch.zeros(M, 2, device='cuda')
        ref_rm.march_rays(n_alive, n_step, al, t, o, d, BOUND, 0.0, 1024, CASCADE, H, scene['bits'], nears, fars, x, dd, dl, 0)
        return x, dd, dl

    def ref_compact(n_alive, a_new, a_old, t_new, t_old, counter):
        # the reference compaction order is nondeterministic; sort to the order-preserving outcome
        ref_rm.compact_rays(n_alive, a_new, a_old, t_new, t_old, counter)
        k = int(counter.item())
        order = torch.argsort(a_new[:k].long(), stable=True)
        a_new[:k] = a_new[:k][order]; t_new[:k] = t_new[:k][order]

    mine = loop(my_march, rm.composite_rays, rm.compact_rays)
    ref = loop(ref_march, ref_rm.composite_rays, ref_compact)
    assert len(mine[3]) == len(ref[3]) and len(ref[3]) > 3
    for (na, ns, x, dl), (rna, rns, rx, rdl) in zip(mine[3], ref[3]):
        assert (na, ns) == (rna, rns)
        assert torch.equal(x.view(torch.int32), rx.view(torch.int32))
        assert torch.equal(dl.view(torch.int32), rdl.view(torch.int32))
    # accumulated outputs: same arithmetic, same order -> tight tolerance (expf vs __expf variants)
    for a, b in zip(mine[:3], ref[:3]):
        assert torch.allclose(a, b, atol=2e-6, rtol=1e-5)
    # CPU restatement of one marching iteration
    x0, _, dl0 = ngp.march_rays(N, 1, np.arange(N, dtype=np.int32), nears.cpu().numpy(), scene['o'][:2048],
                                scene['d'][:2048], BOUND, scene['bits'].cpu().numpy(), CASCADE, H,
                                nears.cpu().numpy(), fars.cpu().numpy())
    assert np.array_equal(x0.view(np.int32), mine[3][0][2][:N].cpu().numpy().view(np.int32))
    assert np.array_equal(dl0.view(np.int32), mine[3][0][3][:N].cpu().numpy().view(np.int32))


def test_mark_untrained_grid_kernel_vs_reference_formulation():
    """al_mark_untrained_grid == the reference's five-loop torch formulation (renderer.py:479-561); the only cells
    allowed to differ sit on a frustum boundary (different fp32 summation order in the 3x3 product)."""
    from autolabel_b200.models import ALNetwork
    from scene_synth import SyntheticScene
    scene = SyntheticScene(12, 48, 64, 16, n_classes=2, seed=3, device='cuda')
    m = ALNetwork(encoding='freq', num_layers=2, hidden_dim=64, num_layers_color=2, hidden_dim_color=64,
                  hidden_dim_semantic=64, semantic_classes=2, bound=3.0, cuda_ray=True).cuda()
    m.density_grid.fill_(0.5)
    want = m.mark_untrained_grid_torch(scene.poses, scene.intrinsics)
    m.mark_untrained_grid(scene.poses, scene.intrinsics)
    got = m.density_grid < 0
    assert 0.05 < want.float().mean().item() < 0.95          # the scene leaves both seen and unseen cells
    mismatch = (got != want).float().mean().item()
    assert mismatch < 1e-4, mismatch
    assert float(m.density_grid[~got].min()) == 0.5           # seen cells untouched


def test_pose_gradients_through_the_operators(scene):
    """near_far_from_aabb and march_rays_train are differentiable w.r.t. the rays like the reference fork's operators
    (raymarching.py:81-136, :358-392): gradients through the autograd Functions equal the closed forms evaluated on the
    operators' own outputs (indices, segments, ts), and rays without grad keep the plain path."""
    from autolabel_b200 import raymarching as rm
    from autolabel_b200.raymarching import _march_train_raw
    o, d = _dev(scene['o']), _dev(scene['d'])
    aabb = _dev(aabb_of(BOUND))
    N = o.shape[0]
    g = torch.Generator().manual_seed(3)
    gn, gf = torch.randn(N, generator=g).cuda(), torch.randn(N, generator=g).cuda()
    o1, d1 = o.clone().requires_grad_(True), d.clone().requires_grad_(True)
    nears, fars, ni, fi = rm.near_far_from_aabb(o1, d1, aabb, 0.2, return_indices=True)
    n0, f0 = rm.near_far_from_aabb(o, d, aabb, 0.2)
    assert torch.equal(nears.detach(), n0) and torch.equal(fars.detach(), f0) and not n0.requires_grad
    hit = (ni != 255)
    (nears[hit] * gn[hit]).sum().backward(retain_graph=True)
    (fars[hit] * gf[hit]).sum().backward()
    zero = torch.zeros_like(gn)
    g_o, g_d = rm.near_far_backward(aabb, o, d, ni, fi, torch.where(hit, gn, zero), torch.where(hit, gf, zero))
    assert torch.allclose(o1.grad, g_o, rtol=1e-6, atol=1e-7) and torch.allclose(d1.grad, g_d, rtol=1e-6, atol=1e-6)
    assert float(o1.grad.abs().sum()) > 0

    o2, d2 = o.clone().requires_grad_(True), d.clone().requires_grad_(True)
    xyzs, dirs, deltas, rays = rm.march_rays_train(o2, d2, BOUND, scene['bits'], CASCADE, H, n0, f0, None, -1, True, 128,
                                                   False, 0, 1024)
    raw = _march_train_raw(o, d, BOUND, scene['bits'], CASCADE, H, n0, f0, None, N * 1024, True, 0.0, 1024)
    m = xyzs.shape[0]
    assert torch.equal(xyzs.detach(), raw['xyzs'][:m]) and torch.equal(rays, raw['rays'])
    G1, G2 = torch.randn(m, 3, generator=g).cuda(), torch.randn(m, 3, generator=g).cuda()
    ((xyzs * G1).sum() + (dirs * G2).sum()).backward()
    pad = torch.zeros(N * 1024 - m, 3, device='cuda')
    g_o, g_d = rm.march_backward(raw['rays'], raw['ts'], torch.cat([G1, pad]), torch.cat([G2, pad]), N)
    # index_add_ accumulates with atomics: the fp32 summation order differs from call to call
    assert torch.allclose(o2.grad, g_o, rtol=1e-4, atol=1e-4) and torch.allclose(d2.grad, g_d, rtol=1e-4, atol=1e-3)
    # segment sums against a direct per-ray loop on a few rays
    r = raw['rays'].cpu().numpy()
    for n in (0, 7, N - 1):
        a, c = int(r[n, 1]), int(r[n, 2])
        if c and a + c <= m:
            assert torch.allclose(o2.grad[r[n, 0]], G1[a:a + c].sum(0), rtol=1e-4, atol=1e-4)
