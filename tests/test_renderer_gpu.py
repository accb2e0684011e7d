"""NeRFRenderer host logic on the sm_100a kernels:
  * `run()` — the path the reference executes today (uniform samples + PyTorch compositing, renderer.py:186-320),
    config C1 shapes (freq encoding, 2x64 MLPs) and the hash-grid model — against the fp32 port oracle/run_path.py;
  * `update_extra_state()` — the occupancy refresh (renderer.py:563-683) reorganised around ONE host read-back
    (occupied-cell counts up front, nonzero_static) — against a literal restatement of the reference's control flow
    (torch.nonzero per cascade, .item() for mean_count) with the same density function and the same RNG seed:
    density grid, bitfield, mean density and mean_count must be IDENTICAL."""
import copy
from types import SimpleNamespace

import pytest
import torch

from tests.helpers import make_rays, record

pytestmark = pytest.mark.gpu


def _model(encoding, hidden, cuda_ray, bound=3.0):
    from autolabel_b200.models import ALNetwork
    torch.manual_seed(0)
    m = ALNetwork(encoding=encoding, num_layers=2, hidden_dim=hidden, num_layers_color=2, hidden_dim_color=hidden,
                  hidden_dim_semantic=64, semantic_classes=2, bound=bound, cuda_ray=cuda_ray).cuda()
    if m._table() is not None:
        with torch.no_grad():
            m._table().uniform_(-0.3, 0.3)
    return m


@pytest.mark.parametrize("encoding,hidden", [("freq", 64), ("hg+freq", 128)])
def test_run_path_matches_port(encoding, hidden):
    from oracle import run_path
    from tests.test_field_gpu import _oracle_inputs
    m = _model(encoding, hidden, cuda_ray=False)
    m.eval()
    N, T = 96, 64
    o, d = make_rays(N, m.bound, seed=3)
    o, d = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    norms = (torch.rand(N, generator=torch.Generator().manual_seed(4)) * 0.3 + 1.0).cuda()
    with torch.no_grad():
        out = m.run(o, d, norms, num_steps=T, perturb=False)
        P, cfg = _oracle_inputs(m)
        field = SimpleNamespace(P=P, cfg=cfg, bound=float(m.bound), min_near=m.min_near, density_scale=m.density_scale)
        ref = run_path.run(field, o, d, norms, num_steps=T, perturb=False)
    assert set(out) == set(ref)
    for k, tol in [('image', 1e-3), ('depth', 1e-3), ('semantic', 1e-3), ('semantic_features', 2e-3),
                   ('coordinates_map', 1e-3), ('depth_variance', 2e-3)]:
        err = (out[k].reshape(ref[k].shape) - ref[k]).abs().max().item()
        if k == 'depth_variance':               # a second moment in squared metres: bound relative to its magnitude
            tol *= max(1.0, ref[k].abs().max().item())
        assert err < tol, f"{k}: {err}"


def _reference_refresh(m, decay=0.95):
    """renderer.py:563-683 with the reference's own synchronisation points (nonzero per cascade, .item())."""
    from autolabel_b200 import raymarching as rm
    H, dev = m.grid_size, m.density_grid.device
    tmp = -torch.ones_like(m.density_grid)

    def query(cas, coords, indices):
        xyzs = 2 * coords.float() / (H - 1) - 1
        bound = min(2 ** cas, m.bound)
        half = bound / H
        pts = xyzs * (bound - half)
        pts += (torch.rand_like(pts) * 2 - 1) * half
        tmp[cas, indices] = m.density_only(pts) * m.density_scale

    if m.iter_density < 16:
        ar = torch.arange(H, dtype=torch.int32, device=dev)
        xx, yy, zz = torch.meshgrid(ar, ar, ar, indexing='ij')
        coords = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=-1)
        indices = rm.morton3D(coords).long()
        for cas in range(m.cascade):
            query(cas, coords, indices)
    else:
        n = H ** 3 // 4
        for cas in range(m.cascade):
            coords = torch.randint(0, H, (n, 3), device=dev)
            indices = rm.morton3D(coords).long()
            occ = torch.nonzero(m.density_grid[cas] > 0).squeeze(-1)
            occ = occ[torch.randint(0, occ.shape[0], [n], dtype=torch.long, device=dev)]
            query(cas, torch.cat([coords, rm.morton3D_invert(occ)], dim=0), torch.cat([indices, occ], dim=0))
    valid = (m.density_grid >= 0) & (tmp >= 0)
    m.density_grid[valid] = torch.maximum(m.density_grid[valid] * decay, tmp[valid])
    mean = torch.mean(m.density_grid.clamp(min=0)).item()
    m.iter_density += 1
    m.density_bitfield.copy_(rm.packbits(m.density_grid, min(mean, m.density_thresh)))
    total = min(16, m.local_step)
    if total > 0:
        m.mean_count = int(m.step_counter[:total, 0].sum().item() / total)
    m.local_step = 0
    return mean


@pytest.mark.parametrize("start_iter", [0, 16])
def test_occupancy_refresh_matches_reference_control_flow(start_iter):
    a = _model("hg+freq", 128, cuda_ray=True)
    with torch.no_grad():                       # a grid with empty, occupied and unseen cells, some recorded steps
        g = torch.Generator().manual_seed(1)
        grid = torch.rand(a.density_grid.shape, generator=g)
        grid = torch.where(grid < 0.3, torch.zeros(()), grid * 5.0)
        grid[:, ::7] = -1.0
        a.density_grid.copy_(grid.cuda())
        a.step_counter[:, 0] = torch.arange(16, dtype=torch.int32, device='cuda') * 1000 + 500000
    a.local_step = 11
    a.iter_density = start_iter
    b = copy.deepcopy(a)
    # After the first 16 refreshes the query set holds duplicate cells (random coordinates + occupied cells drawn
    # with replacement, renderer.py:632-637) and `tmp_grid[cas, indices] = sigmas` keeps whichever duplicate is written
    # last -- order-dependent in the reference too.  Deterministic index_put makes both sides pick the same one.
    torch.use_deterministic_algorithms(True, warn_only=True)
    try:
        _compare_refreshes(a, b)
    finally:
        torch.use_deterministic_algorithms(False)


def _compare_refreshes(a, b):
    for rep in range(2):
        torch.manual_seed(123 + rep)
        a.update_extra_state()
        torch.manual_seed(123 + rep)
        mean_ref = _reference_refresh(b)
        assert torch.equal(a.density_grid, b.density_grid), f"density grid differs after refresh {rep}"
        assert torch.equal(a.density_bitfield, b.density_bitfield)
        assert a.mean_count == b.mean_count and a.local_step == b.local_step == 0
        assert abs(a.mean_density - mean_ref) <= 1e-4 * max(1.0, abs(mean_ref))   # fp32 sum order (block partials + atomics)
        a.local_step = b.local_step = 5


def test_run_path_matches_reference_golden():
    """Config C1 (freq encoding, 2x64 MLPs) through the product's run() against golden outputs of the reference's own
    unmodified ALNetwork.run (tests/golden/ref_run_path.npz, made by tests/golden/make_golden_run.py on CPU with a
    tiny-cuda-nn shim): the six output maps within the north-star tolerances."""
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_run_path.npz"))
    m = _model("freq", 64, cuda_ray=False, bound=float(g["bound"]))
    assert m.semantic_classes == int(g["n_classes"]) and m.hidden_dim_semantic == int(g["feat_dim"])
    with torch.no_grad():
        for net, key in ((m.sigma_net, "w_sigma"), (m.color_net, "w_color"), (m.semantic_features, "w_semf"),
                         (m.semantic_out, "w_semo")):
            w = torch.from_numpy(g[key]).cuda()
            assert net.params.shape == w.shape
            net.params.copy_(w)
    m.eval()
    o, d, norms = (torch.from_numpy(g[k]).cuda() for k in ("rays_o", "rays_d", "direction_norms"))
    with torch.no_grad():
        out = m.run(o, d, norms, num_steps=int(g["num_steps"]), perturb=False)
    for k, tol in [('image', 1e-3), ('depth', 1e-3), ('semantic', 1e-3), ('semantic_features', 2e-3),
                   ('coordinates_map', 1e-3), ('depth_variance', 2e-3)]:
        ref = torch.from_numpy(g["out_" + k]).cuda()
        err = (out[k].reshape(ref.shape) - ref).abs().max().item()
        if k == 'depth_variance':
            tol *= max(1.0, ref.abs().max().item())
        assert err < tol, f"{k}: {err}"


def test_mark_untrained_grid_matches_reference_golden():
    """al_mark_untrained_grid against the mask the reference's OWN five-loop mark_untrained_grid produced on CPU
    (renderer.py:479-561, tests/golden/make_golden_run.py): only cells on a frustum boundary may differ (fp32
    summation order of the 3x3 product)."""
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_run_path.npz"))
    m = _model("freq", 64, cuda_ray=True, bound=float(g["bound"]))
    assert m.cascade == int(g["mark_cascade"])
    m.density_grid.fill_(0.5)
    m.mark_untrained_grid(g["mark_poses"], g["mark_intrinsics"])
    got = (m.density_grid < 0).reshape(-1).cpu().numpy()
    want = np.unpackbits(g["mark_unseen_bits"])[:got.size].astype(bool)
    assert 0.05 < want.mean() < 0.95
    mismatch = float((got != want).mean())
    assert mismatch < 1e-4, mismatch
    assert float(m.density_grid[m.density_grid >= 0].min()) == 0.5     # seen cells untouched


def test_staged_render_matches_reference_golden():
    """render(staged=True) (renderer.py:685-744) on the run() path against the reference's own code: chunking by
    max_ray_batch, and the direction_norms slice -- per-ray norms [B, N] and the [B*N, 1] form of `_get_test`, where
    the reference's slice hands ONE norm (flat pixel b) to the whole row; same behaviour at the API."""
    import os
    import numpy as np
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_run_path.npz"))
    m = _model("freq", 64, cuda_ray=False, bound=float(g["bound"]))
    with torch.no_grad():
        for net, key in ((m.sigma_net, "w_sigma"), (m.color_net, "w_color"), (m.semantic_features, "w_semf"),
                         (m.semantic_out, "w_semo")):
            net.params.copy_(torch.from_numpy(g[key]).cuda())
    m.eval()
    B, N = g["staged2d_depth"].shape
    o = torch.from_numpy(g["rays_o"][:B * N]).cuda().view(B, N, 3)
    d = torch.from_numpy(g["rays_d"][:B * N]).cuda().view(B, N, 3)
    norms = torch.from_numpy(g["direction_norms"][:B * N]).cuda()
    T = int(g["num_steps"])
    with torch.no_grad():
        a = m.render(o, d, norms.view(B, N), staged=True, max_ray_batch=8, num_steps=T, perturb=False)
        b = m.render(o, d, norms.view(-1, 1), staged=True, max_ray_batch=4096, num_steps=T, perturb=False)
    for tag, out in (("staged2d_", a), ("stagedflat_", b)):
        for k, tol in (("image", 1e-3), ("depth", 1e-3), ("semantic_features", 2e-3)):
            ref = torch.from_numpy(g[tag + k]).cuda()
            assert out[k].shape == ref.shape
            assert (out[k] - ref).abs().max().item() < tol, (tag, k)
    assert (a['depth'] - b['depth']).abs().max().item() > 1e-3      # the two norm layouts really differ


def test_update_extra_state_matches_reference_golden():
    """a4 against the reference's OWN NeRFRenderer.update_extra_state (renderer.py:563-683) run unmodified on CPU
    (tests/golden/make_golden_refresh.py): a full refresh and a partial refresh on a grid marked by mark_untrained_grid,
    with the same density field (exact in fp32) and the same random draws (tests/helpers.TorchRngTape replays a numpy
    stream by call order and shape, so any difference in the order, number or shape of the draws changes the result).
    Compared: fp64 sums of the density grid over blocks of 4096 Morton-ordered cells, the bitfield, mean_count /
    iter_density / local_step.  torch's CPU and CUDA kernels round `2 * c / (H - 1)` differently (true division vs
    reciprocal multiply), which moves a handful of the 4.2 M jittered queries across a checker boundary: a few blocks
    may differ by a few checker steps (4.0 each), everything else is bit-identical."""
    import os
    import numpy as np
    from autolabel_b200 import renderer as R
    from tests.helpers import TorchRngTape, checker_density
    here = os.path.dirname(os.path.abspath(__file__))
    g = np.load(os.path.join(here, "golden", "ref_update_extra_state.npz"))
    gm = np.load(os.path.join(here, "golden", "ref_run_path.npz"))
    m = _model("freq", 64, cuda_ray=True, bound=float(g["bound"]))
    m.density_thresh = float(g["density_thresh"])
    shift = [0.0]
    m.density_only = lambda x: checker_density(x + shift[0])
    unseen = np.unpackbits(gm["mark_unseen_bits"])[:m.density_grid.numel()].astype(bool)
    m.density_grid.zero_()
    m.density_grid.view(-1)[torch.from_numpy(unseen).cuda()] = -1.0       # the reference's own mark_untrained_grid mask
    counts = torch.from_numpy(g["step_counts"]).cuda()
    tape = TorchRngTape(torch, int(g["seed"]))
    R.torch = tape
    try:
        for stage, iter_density, sh in (("full", 0, 0.0), ("partial", 16, 0.125)):
            m.iter_density = iter_density
            shift[0] = sh
            m.step_counter.zero_()
            m.step_counter[:counts.numel(), 0] = counts
            m.local_step = int(counts.numel())
            m.update_extra_state()
            grid = m.density_grid.cpu().numpy()
            blocks = grid.astype(np.float64).reshape(-1, 4096).sum(axis=1)
            want = g[stage + "_block_sums"]
            diff = np.abs(blocks - want)
            n_bad = int((diff > 0).sum())
            record("update_extra_state_" + stage, blocks=int(blocks.size), blocks_differing=n_bad, max_block_diff=float(diff.max()))
            assert n_bad <= 0.03 * blocks.size, f"{stage}: {n_bad} of {blocks.size} block sums differ"
            assert diff.max() <= 64.0, f"{stage}: a block differs by {diff.max()}"           # a few checker steps at most
            assert abs(int((grid > 0).sum()) - int(g[stage + "_occupied"])) <= 64, stage
            bits = np.unpackbits(m.density_bitfield.cpu().numpy()) != np.unpackbits(g[stage + "_bitfield"])
            assert bits.mean() < 2e-5, f"{stage}: {int(bits.sum())} bitfield bits differ"
            assert m.mean_count == int(g[stage + "_mean_count"])
            assert m.iter_density == int(g[stage + "_iter_density"]) and m.local_step == int(g[stage + "_local_step"])
            assert abs(m.mean_density - float(g[stage + "_mean_density"])) < 1e-4 * float(g[stage + "_mean_density"])
    finally:
        R.torch = torch
    assert len(tape.calls) == int(g["n_rng_calls"])      # same number of random draws as the reference made
