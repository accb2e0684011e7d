"""Measured baselines of BASELINE.md section 2, taken on the GPU box next to the parity tests (a test may use oracle/):

  B2  the reference's OWN kernels (oracle/_ref, compiled unmodified for sm_100a) timed per kernel against ours on
      identical inputs at C2 sizes: near_far_from_aabb, packbits, march_rays_train, composite_rays_train fwd/bwd
      (3 channels, the only form the reference has), grid_encode fwd/bwd;
  B1  the reference-shaped GPU number: the `run()` path (uniform 256 samples/ray, torch compositing, autograd, torch
      Adam) of oracle/run_path.py on the B200 in fp32 torch -- the stand-in for the reference + tiny-cuda-nn, which
      cannot be installed here.

Results land in gpurun_out/baselines.json (copied to profiles/ per round).  The assertions only check that every
timing was taken and that outputs agree; the numbers are evidence, not gates.
"""
import json
import os

import numpy as np
import pytest
import torch

from tests.helpers import aabb_of, make_density_grid, make_rays

pytestmark = pytest.mark.gpu

BOUND, CASCADE, H = 3.0, 3, 128
_OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "baselines.json")


def _time(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def _save(section, values):
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    data = json.load(open(_OUT)) if os.path.exists(_OUT) else {}
    data[section] = values
    json.dump(data, open(_OUT, "w"), indent=1, sort_keys=True)


def test_b2_reference_kernels_vs_ours(ref_rm, ref_ge):
    from autolabel_b200 import raymarching as rm
    from autolabel_b200.gridencoder import grid_encode
    from autolabel_b200.raymarching import _march_train_raw
    from autolabel_b200 import _lib
    from autolabel_b200._lib import call, ptr, stream_ptr
    from oracle import ngp
    dev = 'cuda'
    st = stream_ptr(torch.device('cuda', torch.cuda.current_device()))
    N = 4096
    o_np, d_np = make_rays(N, BOUND, seed=11)
    o, d = torch.from_numpy(o_np).to(dev), torch.from_numpy(d_np).to(dev)
    grid = torch.from_numpy(make_density_grid(CASCADE, H, seed=12, fill=0.05)).to(dev)
    aabb = torch.from_numpy(aabb_of(BOUND)).to(dev)
    res = {}

    # ---- near_far_from_aabb
    rn, rf = torch.empty(N, device=dev), torch.empty(N, device=dev)
    rni, rfi = torch.empty(N, dtype=torch.uint8, device=dev), torch.empty(N, dtype=torch.uint8, device=dev)
    # `ours_ms`: the C-ABI entry point on preallocated buffers (what the reference's raw pybind call is);
    # `ours_wrapper_ms`: through the Python operator (allocates its outputs, like the reference's own wrapper)
    on, of_ = torch.empty(N, device=dev), torch.empty(N, device=dev)
    oni, ofi = torch.empty(N, dtype=torch.uint8, device=dev), torch.empty(N, dtype=torch.uint8, device=dev)
    res['near_far_from_aabb'] = {
        'ref_ms': _time(lambda: ref_rm.near_far_from_aabb(o, d, aabb, N, 0.2, rn, rf, rni, rfi)),
        'ours_ms': _time(lambda: call("al_near_far_from_aabb", ptr(o), ptr(d), ptr(aabb), N, 0.2, ptr(on), ptr(of_),
                                      ptr(oni), ptr(ofi), st)),
        'ours_wrapper_ms': _time(lambda: rm.near_far_from_aabb(o, d, aabb, 0.2)), 'units': f'{N} rays'}
    nears, fars = rm.near_far_from_aabb(o, d, aabb, 0.2)
    assert torch.equal(on, rn) and torch.equal(of_, rf)

    # ---- packbits
    bits = rm.packbits(grid, 0.01)
    rbits = torch.empty_like(bits)
    obits = torch.empty_like(bits)
    res['packbits'] = {'ref_ms': _time(lambda: ref_rm.packbits(grid, bits.numel(), 0.01, rbits)),
                       'ours_ms': _time(lambda: call("al_packbits", ptr(grid), bits.numel(), 0.01, None, ptr(obits), st)),
                       'ours_wrapper_ms': _time(lambda: rm.packbits(grid, 0.01)), 'units': f'{grid.numel()} cells'}
    assert torch.equal(bits, rbits)

    # ---- march_rays_train (the reference wrapper zero-fills its M x 9 outputs per call, raymarching.py:329-334:
    # timed both with and without that fill)
    M = N * 1024
    xyzs = torch.zeros(M, 3, device=dev); dirs = torch.zeros(M, 3, device=dev)
    deltas = torch.zeros(M, 2, device=dev); ts = torch.zeros(M, 1, device=dev)
    rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)

    def ref_march(fill):
        if fill:
            xyzs.zero_(); dirs.zero_(); deltas.zero_()
        counter.zero_()
        ref_rm.march_rays_train(o, d, bits, BOUND, 0.0, 1024, N, CASCADE, H, M, nears, fars, xyzs, dirs, deltas, ts,
                                rays, counter, 1)

    def our_march():
        return _march_train_raw(o, d, BOUND, bits, CASCADE, H, nears, fars, None, M, True, 0.0, 1024,
                                want_tpos=True, want_sray=True)
    r = our_march()
    mws = torch.empty(_lib.lib.al_march_rays_train_workspace(N, 1024), dtype=torch.uint8, device=dev)
    oxyzs, odirs = torch.empty(M, 3, device=dev), torch.empty(M, 3, device=dev)
    odeltas, ots = torch.empty(M, 2, device=dev), torch.empty(M, 1, device=dev)
    orays, ocounter = torch.empty(N, 3, dtype=torch.int32, device=dev), torch.zeros(2, dtype=torch.int32, device=dev)
    ometa = torch.zeros(2, dtype=torch.int32, device=dev)

    def our_march_raw():       # same outputs as the reference call: xyzs, dirs, deltas, ts, rays, counter
        ocounter.zero_()
        call("al_march_rays_train", ptr(o), ptr(d), ptr(bits), BOUND, 0.0, 1024, N, CASCADE, H, M, ptr(nears), ptr(fars),
             None, 0.2, None, None, ptr(oxyzs), ptr(odirs), ptr(odeltas), ptr(ots), None, None, ptr(orays), ptr(ocounter),
             ptr(ometa), 1, ptr(mws), st)
    our_march_raw()
    ref_march(False)
    torch.cuda.synchronize()
    total = int(counter[0])
    assert total == int(r['counter'][0]) and total > 100000
    res['march_rays_train'] = {'ref_ms': _time(lambda: ref_march(False)), 'ref_with_wrapper_fill_ms': _time(lambda: ref_march(True)),
                               'ours_ms': _time(our_march_raw), 'ours_wrapper_ms': _time(our_march),
                               'units': f'{N} rays, {total} samples'}
    assert torch.equal(orays, r['rays']) and torch.equal(oxyzs[:total], r['xyzs'][:total])

    # ---- composite_rays_train, 3 channels (the form the reference has), on the marched segments
    g = torch.Generator().manual_seed(5)
    sig = (torch.rand(M, generator=g) * 2).to(dev)
    rgb = torch.rand(M, 3, generator=g).to(dev)
    rws, rdepth, rimage = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, 3, device=dev)
    rrays, rdeltas = rays.clone(), deltas.clone()
    ows, odepth, oimage = torch.empty(N, device=dev), torch.empty(N, device=dev), torch.empty(N, 3, device=dev)

    def our_comp_fwd():
        call("al_composite_train_fwd", ptr(sig), 1, ptr(rgb), 3, 3, ptr(r['deltas']), None, None, ptr(r['rays']), M, N, 1.0,
             ptr(ows), ptr(odepth), None, ptr(oimage), None, st)
    res['composite_rays_train_forward_3ch'] = {
        'ref_ms': _time(lambda: ref_rm.composite_rays_train_forward(sig, rgb, rdeltas, rrays, M, N, rws, rdepth, rimage)),
        'ours_ms': _time(our_comp_fwd),
        'ours_wrapper_ms': _time(lambda: rm.composite_rays_train(sig, rgb, r['deltas'], r['rays'])), 'units': f'{total} samples'}
    g_ws, g_img = torch.randn(N, device=dev), torch.randn(N, 3, device=dev)
    rgs, rgr = torch.zeros(M, device=dev), torch.zeros(M, 3, device=dev)
    s1, c1 = sig.clone().requires_grad_(True), rgb.clone().requires_grad_(True)
    ws, depth, image = rm.composite_rays_train(s1, c1, r['deltas'], r['rays'])
    lossv = (ws * g_ws).sum() + (image * g_img).sum()
    ogs, ogr = torch.zeros(M, device=dev), torch.zeros(M, 3, device=dev)
    our_comp_fwd()

    def our_comp_bwd():        # no depth gradient, like the reference (raymarching.py:437-438)
        call("al_composite_train_bwd", ptr(g_ws), None, ptr(g_img), ptr(sig), 1, ptr(rgb), 3, 3, ptr(r['deltas']), None,
             ptr(r['rays']), ptr(ows), ptr(odepth), ptr(oimage), M, N, 1.0, ptr(ogs), 1, ptr(ogr), 3, None, st)
    res['composite_rays_train_backward_3ch'] = {
        'ref_ms': _time(lambda: ref_rm.composite_rays_train_backward(g_ws, g_img, sig, rgb, rdeltas, rrays, rws, rimage, M, N, rgs, rgr)),
        'ours_ms': _time(our_comp_bwd),
        'ours_wrapper_ms': _time(lambda: torch.autograd.grad(lossv, (s1, c1), retain_graph=True)), 'units': f'{total} samples'}
    gs_w, gr_w = torch.autograd.grad(lossv, (s1, c1), retain_graph=True)
    assert torch.equal(ogs[:total], gs_w[:total]) and torch.equal(ogr[:total], gr_w[:total])

    # ---- hash grid (hg+freq hyper-parameters) on B samples inside [0,1]^3
    B, L, C = 1 << 20, 16, 2
    offsets = torch.from_numpy(ngp.grid_offsets(16, 16, 2.0, 19, 3)).to(dev)
    x = torch.rand(B, 3, generator=g).to(dev)
    table = ((torch.rand(int(offsets[-1]), C, generator=g) * 2 - 1) * 0.1).to(dev)
    rout = torch.empty(L, B, C, device=dev)
    oout = torch.empty(L, B, C, device=dev)
    og = torch.zeros_like(table)
    dummy = torch.empty(1, device=dev)
    emb = table.clone().requires_grad_(True)
    out = grid_encode(x, emb, offsets, 2.0, 16, False, 0)
    gout = torch.randn(B, L * C, device=dev)
    gl = gout.view(B, L, C).permute(1, 0, 2).contiguous()
    rg = torch.zeros_like(table)
    res['grid_encode_forward'] = {
        'ref_ms': _time(lambda: ref_ge.grid_encode_forward(x, table, offsets, rout, B, 3, C, L, 1.0, 16, False, dummy, 0), iters=10),
        'ours_ms': _time(lambda: call("al_grid_encode_forward", ptr(x), ptr(table), ptr(offsets), ptr(oout), B, 3, C, L, 1.0, 16,
                                      0, None, 0, None, st), iters=10),
        'ours_wrapper_ms': _time(lambda: grid_encode(x, table, offsets, 2.0, 16, False, 0), iters=10), 'units': f'{B} samples'}
    assert torch.equal(oout, rout)

    def ref_grid_bwd():
        # grid.py:70 permutes the incoming gradient to level-major before the kernel
        glm = gout.view(B, L, C).permute(1, 0, 2).contiguous()
        ref_ge.grid_encode_backward(glm, x, table, offsets, rg, B, 3, C, L, 1.0, 16, False, dummy, dummy, 0)
    res['grid_encode_backward'] = {
        'ref_ms': _time(ref_grid_bwd, iters=10),
        'ours_ms': _time(lambda: call("al_grid_encode_backward", ptr(gl), ptr(x), ptr(offsets), ptr(og), B, 3, C, L, 1.0, 16, 0,
                                      None, None, 0, st), iters=10),
        'ours_with_permute_ms': _time(lambda: call("al_grid_encode_backward",
                                                   ptr(gout.view(B, L, C).permute(1, 0, 2).contiguous()), ptr(x), ptr(offsets),
                                                   ptr(og), B, 3, C, L, 1.0, 16, 0, None, None, 0, st), iters=10),
        'ours_wrapper_ms': _time(lambda: torch.autograd.grad(out, emb, gout, retain_graph=True), iters=10), 'units': f'{B} samples'}
    res['grid_encode_backward']['ref_ms_note'] = 'reference timing includes the level-major permute of grid.py:70; ours_with_permute_ms likewise'
    for k, v in res.items():
        v['speedup'] = v['ref_ms'] / v['ours_ms']
        v['wrapper_speedup'] = v['ref_ms'] / v['ours_wrapper_ms']
        assert v['ref_ms'] > 0 and v['ours_ms'] > 0, k
    _save('B2_reference_kernels_sm100a', res)


def test_b1_reference_shaped_run_path_on_gpu():
    """oracle/run_path.py (port of renderer.run + train_step + torch Adam) on the B200, fp32, C2 shapes:
    4096 rays x 256 uniform samples.  Chunked over rays only if the [N, T, F] autograd intermediates do not fit."""
    from oracle import run_path
    dev = 'cuda'
    N, F = 4096, 64
    field = run_path.OracleField('hg+freq', 128, 128, F, 2, bound=BOUND, seed=0, device=dev)
    opt = field.optimizer()
    g = torch.Generator().manual_seed(0)
    dd = torch.randn(N, 3, generator=g)
    data = {'rays_o': ((torch.rand(N, 3, generator=g) - 0.5) * 2.0).to(dev), 'rays_d': (dd / dd.norm(dim=1, keepdim=True)).to(dev),
            'direction_norms': torch.ones(N, 1, device=dev), 'pixels': torch.rand(N, 3, generator=g).to(dev),
            'depth': (torch.rand(N, generator=g) * 3).to(dev), 'semantic': torch.randint(-1, 2, (N,), generator=g).to(dev),
            'features': torch.rand(N, F, generator=g).to(dev)}
    losses = []

    def step():
        losses.append(run_path.train_step(field, opt, data))
    ms = _time(step, iters=5, warm=2)
    assert np.isfinite(losses[-1])
    _save('B1_run_path_torch_fp32_on_b200', {
        'ms_per_step': ms, 'rays_per_s': N / (ms * 1e-3), 'rays': N, 'samples_per_ray': 256,
        'what': 'oracle/run_path.py: uniform sampling, torch fp32 field (hash grid via index ops), torch compositing, autograd, '
                'torch.optim.Adam -- stand-in for the reference run() path with tiny-cuda-nn (not installable here)'})


def test_b1_reference_shaped_run_path_on_gpu_amp():
    """B1 under AMP, the precision the reference trains in (`fp16=True`: autocast + GradScaler, autolabel/trainer.py:40-48):
    the same port with the field evaluated under torch.autocast(float16) (MLP matmuls in fp16 tensor-core GEMMs, as tcnn
    runs them) and a GradScaler around backward / step."""
    from oracle import run_path
    dev = 'cuda'
    N, F = 4096, 64
    field = run_path.OracleField('hg+freq', 128, 128, F, 2, bound=BOUND, seed=0, device=dev)
    opt = field.optimizer()
    scaler = torch.amp.GradScaler('cuda')
    g = torch.Generator().manual_seed(0)
    dd = torch.randn(N, 3, generator=g)
    data = {'rays_o': ((torch.rand(N, 3, generator=g) - 0.5) * 2.0).to(dev), 'rays_d': (dd / dd.norm(dim=1, keepdim=True)).to(dev),
            'direction_norms': torch.ones(N, 1, device=dev), 'pixels': torch.rand(N, 3, generator=g).to(dev),
            'depth': (torch.rand(N, generator=g) * 3).to(dev), 'semantic': torch.randint(-1, 2, (N,), generator=g).to(dev),
            'features': torch.rand(N, F, generator=g).to(dev)}
    losses = []

    def step():
        opt.zero_grad()
        with torch.autocast('cuda', dtype=torch.float16):
            out = run_path.run(field, data['rays_o'], data['rays_d'], data['direction_norms'], num_steps=256, perturb=True)
            loss = run_path.loss_fn({k: v.float() for k, v in out.items()}, data)
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        losses.append(float(loss.item()))
    ms = _time(step, iters=5, warm=2)
    assert np.isfinite(losses[-1])
    _save('B1_run_path_torch_amp_on_b200', {
        'ms_per_step': ms, 'rays_per_s': N / (ms * 1e-3), 'rays': N, 'samples_per_ray': 256,
        'what': 'oracle/run_path.py under torch.autocast(float16) + GradScaler (the reference trains with fp16=True): uniform '
                'sampling, torch field (fp16 matmuls), torch compositing, autograd, torch.optim.Adam -- stand-in for the '
                'reference run() path with tiny-cuda-nn (not installable here)'})
