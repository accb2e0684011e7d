"""Tier O2 parity of the fused field and of the full marched render (forward + parameter
gradients) against the fp32 PyTorch restatement in oracle/field_oracle.py, at the C1 (freq, 2x64) and
C2 (hg+freq, 128-wide, F=64) model shapes.  Tolerances are BASELINE.json's: 1e-3 absolute for rgb /
depth / logits / gradients, 2e-3 for features."""
import numpy as np
import pytest
import torch

from tests.helpers import aabb_of, make_density_grid, make_rays, record, rel_l2

pytestmark = pytest.mark.gpu


def _model(encoding, hidden, F=64, C=2, bound=3.0, table_scale=0.3, seed=0):
    from autolabel_b200.models import ALNetwork
    torch.manual_seed(seed)
    m = ALNetwork(encoding=encoding, num_layers=2, hidden_dim=hidden, geo_feat_dim=15, num_layers_color=2,
                  hidden_dim_color=hidden, hidden_dim_semantic=F, semantic_classes=C, bound=bound, cuda_ray=True).cuda()
    t = m._table()
    if t is not None:
        with torch.no_grad():
            t.uniform_(-table_scale, table_scale)   # trained-scale features, not the 1e-4 init
    return m


def _oracle_inputs(m):
    from oracle import field_oracle as fo
    offsets, L, S, H = m._grid_meta()
    P = dict(w_sigma=m.sigma_net.params.detach().clone().requires_grad_(True),
             w_color=m.color_net.params.detach().clone().requires_grad_(True),
             w_semf=m.semantic_features.params.detach().clone().requires_grad_(True),
             w_semo=m.semantic_out.params.detach().clone().requires_grad_(True))
    t = m._table()
    if t is not None:
        P['table'] = t.detach().clone().view(-1, 2).requires_grad_(True)
    cfg = dict(encoding=m.encoding, bound=m.bound, hidden=m.hidden_dim, hidden_color=m.hidden_dim_color,
               feat_dim=m.hidden_dim_semantic, n_classes=m.semantic_classes,
               offsets=None if offsets is None else offsets.cpu().numpy(), per_level_scale=float(2.0 ** S), H=H)
    return P, cfg


@pytest.mark.parametrize("encoding,hidden", [("hg+freq", 128), ("freq", 64), ("hg", 128)])
def test_field_forward_vs_oracle(encoding, hidden):
    from oracle import field_oracle as fo
    m = _model(encoding, hidden)
    g = torch.Generator().manual_seed(3)
    n = 5000
    xyz = ((torch.rand(n, 3, generator=g) * 2 - 1) * m.bound).cuda()
    d = torch.randn(n, 3, generator=g)
    d = (d / d.norm(dim=1, keepdim=True)).cuda()
    vals = m.field_values(xyz, d)
    P, cfg = _oracle_inputs(m)
    with torch.no_grad():
        sigma, rgb, logits, feat, h = fo.field_forward(xyz, d, P, cfg)
    C, F = m.semantic_classes, m.hidden_dim_semantic
    # PER-SAMPLE (un-composited) outputs: bounded by fp16 operand rounding through three layers,
    # RAW = 4e-3 on O(1) values; the north-star bars (1e-3 / 2e-3) apply to the RENDERED outputs and are
    # enforced in test_render_train_step_vs_oracle.
    RAW = 4e-3
    smax = max(1.0, sigma.abs().max().item())
    errs = dict(sigma=(vals[:, 0] - sigma).abs().max().item() / smax, rgb=(vals[:, 1:4] - rgb).abs().max().item(),
                logits=(vals[:, 4:4 + C] - logits).abs().max().item(),
                feat=(vals[:, 4 + C:4 + C + F] - feat).abs().max().item())
    record(f"field_forward_{encoding}_{hidden}", **errs)
    assert errs['sigma'] < RAW and errs['rgb'] < 1e-3 and errs['logits'] < RAW and errs['feat'] < RAW, errs
    # module-level API (models.py:175-256) agrees too
    dens = m.density(xyz)
    assert (dens['sigma'] - sigma).abs().max().item() < RAW * smax
    assert (dens['geo_feat'] - h[:, 1:]).abs().max().item() < RAW
    sem, sf = m.semantic(dens['geo_feat'], dens['sigma'])
    assert (sem - logits).abs().max().item() < RAW and (sf - feat).abs().max().item() < RAW
    col = m.color(xyz, d, geo_feat=dens['geo_feat'])
    assert (col - rgb).abs().max().item() < 1e-3
    assert (m.density_only(xyz) - sigma).abs().max().item() < RAW * smax


@pytest.fixture(params=["tcgen05", "mma_sync"])
def backend(request):
    """tcgen05: TMEM MLPs + rank-1 compositing backward; mma_sync: the baseline kernels + materialised gradients."""
    from autolabel_b200 import _lib
    prev = _lib.lib.al_set_mlp_backend(1 if request.param == "tcgen05" else 0)
    yield request.param
    _lib.lib.al_set_mlp_backend(prev)


@pytest.mark.parametrize("thresh", [0.0, 1e-4])
@pytest.mark.parametrize("encoding,hidden", [("hg+freq", 128), ("freq", 64)])
def test_render_train_step_vs_oracle(encoding, hidden, backend, thresh):
    """model.render() in training mode == oracle(field on the marched samples + ragged compositing over EVERY marched
    sample); the same loss gives the same parameter gradients.  thresh = 0: every marched sample goes through the
    heads (the reference's training kernels, the default); 1e-4: opt-in training-time early termination."""
    m = _model(encoding, hidden, table_scale=0.3)
    m.train_t_thresh = thresh
    _check_render_train(m, f"render_train_{backend}_{encoding}_{hidden}_t{thresh:g}")


@pytest.mark.parametrize("density_scale", [50.0, 200.0])
def test_render_train_early_termination_on_opaque_scenes(density_scale):
    """Dense fields (rays saturate after a few samples): most marched samples are cut by the opt-in
    train_t_thresh = 1e-4, outputs and parameter gradients still match the fp32 oracle that composites every sample
    within the north-star tolerances."""
    m = _model("hg+freq", 128, table_scale=0.3)
    m.density_scale = density_scale
    assert m.train_t_thresh == 0.0      # default: exact reference semantics (every marched sample composited)
    m.train_t_thresh = 1e-4
    _check_render_train(m, f"render_train_early_term_scale{density_scale:g}", max_alive_fraction=0.9)


@pytest.mark.parametrize("F,C", [(512, 2), (64, 40), (128, 2), (512, 606)])
def test_render_train_step_wide_heads(F, C):
    """Config C5 (512-d LSeg feature head) and ScanNet-sized label sets through the SAME fused render path: the heads
    that do not fit the weight-resident kernels run as tiled tcgen05 GEMMs (csrc/gemm_tc.cu) inside al_field_*."""
    m = _model("hg+freq", 128, F=F, C=C, table_scale=0.3)
    _check_render_train(m, f"render_train_wide_F{F}_C{C}", N=192)


def _check_render_train(m, tag, N=384, max_alive_fraction=None):
    from autolabel_b200 import raymarching as rm
    from autolabel_b200.raymarching import _march_train_raw
    from oracle import field_oracle as fo
    m.train()
    o, d = make_rays(N, m.bound, seed=5, inside=False)
    o, d = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    grid = torch.from_numpy(make_density_grid(m.cascade, 128, seed=6, fill=0.04)).cuda()
    m.density_grid.copy_(grid)
    m.density_bitfield.copy_(rm.packbits(grid, 0.01))
    norms = (torch.rand(N, 1, generator=torch.Generator().manual_seed(1)) * 0.3 + 1.0).cuda()
    out = m.render(o, d, norms, staged=False, bg_color=None, perturb=True)
    C, F = m.semantic_classes, m.hidden_dim_semantic
    assert set(out) == {'depth', 'depth_variance', 'image', 'semantic', 'semantic_features', 'coordinates_map'}
    assert out['image'].shape == (N, 3) and out['semantic'].shape == (N, C) and out['semantic_features'].shape == (N, F)
    alive, marched = int(m.last_alive_meta[0]), int(m.last_meta[0])
    assert 0 < alive <= marched and (m.train_t_thresh > 0 or alive == marched)
    if max_alive_fraction is not None:
        assert alive < max_alive_fraction * marched, (alive, marched)

    # oracle on the same samples
    nears, fars = rm.near_far_from_aabb(o, d, m.aabb_train, m.min_near)
    M = N * 1024
    r = _march_train_raw(o, d, m.bound, m.density_bitfield, m.cascade, 128, nears, fars, None, M, True, 0.0, 1024,
                         want_tpos=True, want_sray=True)
    tot = int(r['counter'][0])
    assert 1000 < tot < 60000, tot
    P, cfg = _oracle_inputs(m)
    xyz, dirs = r['xyzs'][:tot], d[r['sray'][:tot].long()]
    sigma, rgb, logits, feat, _ = fo.field_forward(xyz, dirs, P, cfg)
    vals = torch.cat([rgb, logits, feat], dim=1)
    ows, odepth, odsq, oout, ocoords = fo.composite(sigma, vals, r['deltas'][:tot], r['tpos'][:tot], xyz, r['rays'], M,
                                                    sigma_scale=m.density_scale)
    ref = fo.render_outputs(ows, odepth, odsq, oout, ocoords, norms.view(-1), C)
    # depth_variance = sum w (t - depth)^2 is a detached second moment in squared metres (renderer.py:277-278), not one of
    # the north-star outputs.  The samples cut by train_t_thresh carry a total weight below the threshold, each with
    # (t - depth)^2 up to the squared chord of the box (2 sqrt(3) bound)^2, so their share of the sum is bounded by
    # thresh * chord^2; every first-moment output stays inside its 1e-3 / 2e-3 bar (thresh * chord < 1.1e-3 * w_sum).
    chord2 = (2.0 * 3.0 ** 0.5 * float(m.bound)) ** 2
    tol = {'image': 1e-3, 'depth': 1e-3, 'semantic': 1e-3, 'semantic_features': 2e-3, 'coordinates_map': 1e-3,
           'depth_variance': 2e-3 + float(m.train_t_thresh) * chord2}
    rep = {}
    for k, t in tol.items():
        err = (out[k] - ref[k]).abs().max().item()
        rep[k] = err
        assert err < t, f"{k}: {err}"

    # identical loss on both sides (the loss of autolabel/trainer.py:54-94 without the data terms' masks)
    g = torch.Generator().manual_seed(2)
    gt_rgb, gt_depth = torch.rand(N, 3, generator=g).cuda(), (torch.rand(N, generator=g) * 4).cuda()
    gt_feat, gt_sem = torch.randn(N, F, generator=g).cuda(), torch.randint(0, C, (N,), generator=g).cuda()

    def loss_of(o_):
        return (((o_['image'] - gt_rgb) ** 2).mean() + 0.1 * (o_['depth'] - gt_depth).abs().mean() +
                0.5 * torch.nn.functional.l1_loss(o_['semantic_features'], gt_feat) +
                torch.nn.functional.cross_entropy(o_['semantic'], gt_sem))

    l1, l2 = loss_of(out), loss_of(ref)
    assert abs(l1.item() - l2.item()) < 1e-3
    scale = 1024.0    # a GradScaler-like loss scale; gradients are compared after unscaling
    (l1 * scale).backward()
    (l2 * scale).backward()
    pairs = [(m.sigma_net.params.grad, P['w_sigma'].grad, 'sigma_net'), (m.color_net.params.grad, P['w_color'].grad, 'color_net'),
             (m.semantic_features.params.grad, P['w_semf'].grad, 'semantic_features'),
             (m.semantic_out.params.grad, P['w_semo'].grad, 'semantic_out')]
    if m._table() is not None:
        pairs.append((m._table().grad.view(-1, 2), P['table'].grad, 'hash table'))
    for a, b, name in pairs:
        a, b = a / scale, b / scale
        err = (a - b).abs().max().item()
        rl2 = rel_l2(a, b)
        rep['grad_abs_' + name] = err
        rep['grad_rel_l2_' + name] = rl2
        assert err < 1e-3, f"{name}: abs {err}"          # north star: parameter gradients within 1e-3 absolute
        assert rl2 < 3e-2, f"{name}: relative L2 {rl2:.3e}"  # fp16 operands + ReLU-boundary flips (see test_mlp_gpu)
    rep['samples'] = tot
    rep['alive_samples'] = alive
    record(tag, **rep)


def test_render_eval_matches_train_march():
    """Inference path (exact budget, chunked) == the training path on the same rays with perturb off."""
    from autolabel_b200 import raymarching as rm
    m = _model("hg+freq", 128)
    N = 1000
    o, d = make_rays(N, m.bound, seed=8)
    o, d = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    grid = torch.from_numpy(make_density_grid(m.cascade, 128, seed=9, fill=0.04)).cuda()
    m.density_grid.copy_(grid)
    m.density_bitfield.copy_(rm.packbits(grid, 0.01))
    norms = torch.ones(N, 1).cuda()
    m.train()
    m.train_t_thresh = 0.0      # composite every marched sample: the exact counterpart of early_termination=False
    with torch.no_grad():
        a = m.render(o, d, norms, perturb=False, force_all_rays=True)
    m.eval()
    m.max_render_rays = 300   # force several chunks
    b = m.render(o.view(1, N, 3), d.view(1, N, 3), norms, staged=True, perturb=False, early_termination=False)
    for k in a:
        assert torch.allclose(a[k].reshape(-1), b[k].reshape(-1), atol=1e-6), k
    # early termination (the default inference path): a ray stops after the sample that starts with T < 1e-4, so
    # every output is within 1e-4 x (value range) of the full composite
    c = m.render(o.view(1, N, 3), d.view(1, N, 3), norms, staged=True, perturb=False)
    assert int(m.last_meta[1]) > 0
    for k, tol in [('image', 3e-4), ('semantic', 1e-3), ('semantic_features', 2e-3), ('coordinates_map', 1e-3), ('depth', 1e-3)]:
        assert (a[k].reshape(-1) - c[k].reshape(-1)).abs().max().item() < tol, k


@pytest.mark.parametrize("wave_steps", [(32,) * 8 + (64, 64, 128, 256, 512), (7, 19, 40, 100, 512)])
def test_render_waves_fused_compositing_matches_value_matrix_path(wave_steps):
    """Inference waves: compositing folded into the head epilogues (al_composite_rays_weights +
    al_field_heads_forward_sum) against the value-matrix path (al_field_forward + al_composite_rays, the literal
    renderer.py:440-460 sequence).  Ray-level sums (depth, weights, coordinates) come from the same arithmetic in the
    same order and must be bit-identical; the channel sums differ by fp32 summation order only.  The second schedule
    has waves that are not multiples of 32 samples, so a warp's 32 rows span several rays."""
    from autolabel_b200 import raymarching as rm
    m = _model("hg+freq", 128)
    if not m.fused_wave_composite():
        pytest.skip("needs the tcgen05 back end")
    N = 1500
    o, d = make_rays(N, m.bound, seed=18)
    o, d = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    grid = torch.from_numpy(make_density_grid(m.cascade, 128, seed=19, fill=0.04)).cuda()
    m.density_grid.copy_(grid)
    m.density_bitfield.copy_(rm.packbits(grid, 0.01))
    norms = torch.ones(N, 1).cuda()
    m.eval()
    m.wave_steps = wave_steps
    res = {}
    for fused in (False, True):
        m.fused_composite = fused
        assert m.fused_wave_composite() == fused
        res[fused] = m.render(o.view(1, N, 3), d.view(1, N, 3), norms, staged=True, perturb=False)
        res[fused] = {k: v.clone() for k, v in res[fused].items()}
    a, b = res[False], res[True]
    for k in ('depth', 'coordinates_map'):
        assert torch.equal(a[k], b[k]), k
    for k, tol in [('image', 2e-6), ('semantic', 2e-5), ('semantic_features', 2e-5)]:
        err = (a[k] - b[k]).abs().max().item()
        assert err < tol * max(1.0, a[k].abs().max().item()), (k, err)
