"""Per-phase / per-kernel CUDA-event timing for bench.py (`roofline` and `phases_ms`).

A real training batch is pushed through the same C-ABI calls the product path makes
(autolabel_b200/renderer.py::_FusedRender), keeping every intermediate buffer; each phase and each
candidate hot kernel is then re-launched standalone `reps` times between two CUDA events on the
launching (current torch) stream.  The kernel with the largest average launch time becomes the
`roofline` entry:  achieved = algorithmic work per launch / average launch duration, against the
measured peaks in MEASURED_PEAKS.json (burst figures: the kernel is timed alone), else the fallback
of /opt/skills/guides/B200_PROFILING.md.  Algorithmic work per sample is stated in DESIGN.md section 4.
"""
import ctypes
import json
import os

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm": float(p["hbm_gbs"]), "tensor": float(p["bf16_tflops"]), "source": "MEASURED_PEAKS.json (burst)"}
    return {"hbm": 6650.0, "tensor": 1590.0, "source": "fallback of B200_PROFILING.md"}


# kernel (as `roofline.kernel` names it) -> regex of its ncu kernel name in profiles/roofline_traffic.json
_NCU_NAME = {"encode_position": "k_encode_position", "sigma_mlp_forward": "k_mlp_fwd_tc<48", "sigma_mlp_backward": "k_mlp_bwd_tc<48",
             "grid_scatter": "k_grid_bwd<", "adam_table": "k_adam"}


def committed_traffic(kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of `kernel` from the committed `ncu --set full`
    capture of one training step (profiles/roofline_traffic.json, written by tools/summarize_ncu.py traffic); None if absent."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(path):
        return None
    rows = json.load(open(path)).get("kernels", {})
    key = _NCU_NAME.get(kernel, kernel)
    hits = [v for k, v in rows.items() if k.startswith(key)]
    if not hits:
        return None
    return max(h["dram_bytes"] for h in hits), json.load(open(path)).get("capture", {}).get("marched_samples")


def _time(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def measure(args, scene, model, trainer, device, n_rays):
    from autolabel_b200 import _lib
    from autolabel_b200._lib import call, ptr, stream_ptr
    lib = _lib.lib
    st = stream_ptr(device)
    batch = scene.next_train(n_rays)
    rays_o, rays_d = batch['rays_o'].contiguous(), batch['rays_d'].contiguous()
    N = rays_o.shape[0]
    K = model.n_channels
    ldv = 1 + K
    F, C = model.hidden_dim_semantic, model.semantic_classes
    desc = model.field_desc()
    dref = ctypes.byref(desc)
    max_steps = 1024
    M = model.mean_count + 128 - model.mean_count % 128 if model.mean_count > 0 else N * max_steps
    dev = device
    f32, i32 = torch.float32, torch.int32
    xyzs = torch.empty(M, 3, dtype=f32, device=dev); deltas = torch.empty(M, 2, dtype=f32, device=dev)
    tpos = torch.empty(M, dtype=f32, device=dev); sray = torch.empty(M, dtype=i32, device=dev)
    rays = torch.empty(N, 3, dtype=i32, device=dev); meta = torch.zeros(2, dtype=i32, device=dev)
    counter = torch.zeros(2, dtype=i32, device=dev)
    mws = torch.empty(lib.al_march_rays_train_workspace(N, max_steps), dtype=torch.uint8, device=dev)
    vals = torch.empty(M, ldv, dtype=f32, device=dev); g_vals = torch.zeros(M, ldv, dtype=f32, device=dev)
    fws = torch.empty(lib.al_field_workspace(dref, M, 1), dtype=torch.uint8, device=dev)
    ws = torch.empty(N, dtype=f32, device=dev); depth = torch.empty(N, dtype=f32, device=dev)
    dsq = torch.empty(N, dtype=f32, device=dev); out = torch.empty(N, K, dtype=f32, device=dev)
    coords = torch.empty(N, 3, dtype=f32, device=dev)
    g_ws = torch.randn(N, device=dev) * 1e-4; g_depth = torch.randn(N, device=dev) * 1e-5
    g_out = torch.randn(N, K, device=dev) * 1e-4
    params = model.field_params()
    amax_t = torch.zeros(1, dtype=f32, device=dev)
    grads = [torch.zeros_like(p) for p in params]

    def march():
        counter.zero_()
        call("al_march_rays_train", ptr(rays_o), ptr(rays_d), ptr(model.density_bitfield), float(model.bound), 0.0,
             max_steps, N, int(model.cascade), int(model.grid_size), int(M), None, None, ptr(model.aabb_train),
             float(model.min_near), None, None, ptr(xyzs), None, ptr(deltas), None, ptr(tpos), ptr(sray), ptr(rays),
             ptr(counter), ptr(meta), 1, ptr(mws), st)

    thr = float(getattr(model, 'train_t_thresh', 0.0))
    early = thr > 0.0
    if early:   # staging buffers of the march; the alive prefixes are packed into the buffers the later phases read
        p_xyzs, p_deltas, p_tpos, p_sray, p_rays, p_meta = xyzs, deltas, tpos, sray, rays, meta
        xyzs = torch.empty(M, 3, dtype=f32, device=dev); deltas = torch.empty(M, 2, dtype=f32, device=dev)
        tpos = torch.empty(M, dtype=f32, device=dev); sray = torch.empty(M, dtype=i32, device=dev)
        rays = torch.empty(N, 3, dtype=i32, device=dev); meta = torch.zeros(2, dtype=i32, device=dev)
        p_xenc = torch.empty(M, int(desc.in_pad), dtype=torch.float16, device=dev)
        p_h16 = torch.empty(M, 16, dtype=f32, device=dev); p_sigma = torch.empty(M, dtype=f32, device=dev)
        alive_ws = torch.empty(N, dtype=i32, device=dev)
        slot_x, slot_h = ctypes.c_void_p(), ctypes.c_void_p()
        call("al_field_workspace_slots", dref, M, 1, ptr(fws), ctypes.byref(slot_x), ctypes.byref(slot_h))

        def march():
            counter.zero_()
            call("al_march_rays_train", ptr(rays_o), ptr(rays_d), ptr(model.density_bitfield), float(model.bound), 0.0,
                 max_steps, N, int(model.cascade), int(model.grid_size), int(M), None, None, ptr(model.aabb_train),
                 float(model.min_near), None, None, ptr(p_xyzs), None, ptr(p_deltas), None, ptr(p_tpos), ptr(p_sray),
                 ptr(p_rays), ptr(counter), ptr(p_meta), 1, ptr(mws), st)

        def density_pre():
            call("al_field_density_pre", dref, ptr(p_xyzs), M, ptr(p_meta), ptr(p_xenc), ptr(p_h16), ptr(p_sigma), st)

        def compact():
            call("al_compact_alive", ptr(p_sigma), ptr(p_deltas), ptr(p_rays), M, N, float(model.density_scale), thr,
                 ptr(p_xyzs), ptr(p_tpos), ptr(p_sray), ptr(p_xenc), int(desc.in_pad), ptr(p_h16), ptr(rays), ptr(meta),
                 ptr(xyzs), ptr(deltas), ptr(tpos), ptr(sray), slot_x, slot_h, ptr(vals), ldv, ptr(alive_ws), st)

        def heads_fwd():
            call("al_field_heads_forward", dref, ptr(rays_d), ptr(sray), M, ptr(meta), ptr(vals), ldv, ptr(fws), st)

        def field_fwd():
            density_pre(); compact(); heads_fwd()
    else:
        def field_fwd():
            call("al_field_forward", dref, ptr(xyzs), ptr(rays_d), ptr(sray), M, ptr(meta), ptr(vals), ldv, None, 0, ptr(fws), st)

    def comp_fwd():
        call("al_composite_train_fwd", ptr(vals), ldv, vals.data_ptr() + 4, ldv, K, ptr(deltas), ptr(tpos), ptr(xyzs),
             ptr(rays), M, N, float(model.density_scale), ptr(ws), ptr(depth), ptr(dsq), ptr(out), ptr(coords), st)

    tc = lib.al_set_mlp_backend(-1) == 1
    w_s = torch.empty(M, dtype=f32, device=dev); g_sig = torch.empty(M, dtype=f32, device=dev)

    def comp_bwd():
        if tc:
            call("al_composite_train_bwd_weights", ptr(g_ws), ptr(g_depth), ptr(g_out), ptr(vals), ldv, vals.data_ptr() + 4,
                 ldv, K, ptr(deltas), ptr(tpos), ptr(rays), ptr(ws), ptr(depth), ptr(out), M, N, float(model.density_scale),
                 ptr(w_s), ptr(g_sig), ptr(amax_t), st)
        else:
            call("al_composite_train_bwd", ptr(g_ws), ptr(g_depth), ptr(g_out), ptr(vals), ldv, vals.data_ptr() + 4, ldv, K,
                 ptr(deltas), ptr(tpos), ptr(rays), ptr(ws), ptr(depth), ptr(out), M, N, float(model.density_scale),
                 ptr(g_vals), ldv, g_vals.data_ptr() + 4, ldv, ptr(amax_t), st)

    def field_bwd():
        if tc:
            call("al_field_backward_rays", dref, ptr(xyzs), M, ptr(meta), ptr(vals), ldv, ptr(w_s), ptr(g_sig), ptr(g_out),
                 ptr(sray), ptr(amax_t), ptr(grads[0]), ptr(grads[1]), ptr(grads[2]), ptr(grads[3]), ptr(grads[4]), ptr(fws), st)
        else:
            call("al_field_backward", dref, ptr(xyzs), M, ptr(meta), ptr(vals), ptr(g_vals), ptr(amax_t), ldv, ptr(grads[0]),
                 ptr(grads[1]), ptr(grads[2]), ptr(grads[3]), ptr(grads[4]), ptr(fws), st)

    adam_state = [(torch.zeros_like(p), torch.zeros_like(p)) for p in params]
    shadow = [p.detach().clone() for p in params]       # do not disturb the trained parameters

    def adam():
        for p, g, (m, v) in zip(shadow, grads, adam_state):
            call("al_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), 5e-3, 0.9, 0.99, 1e-15, 0.0, 10, 1.0, 1, st)

    march(); field_fwd(); comp_fwd(); comp_bwd(); field_bwd()
    torch.cuda.synchronize()
    n_live = int(meta[0].item())                       # samples the heads / compositing / backward run on
    n_marched = int(p_meta[0].item()) if early else n_live

    phases = {"march": _time(march), "field_forward": _time(field_fwd), "composite_forward": _time(comp_fwd),
              "composite_backward": _time(comp_bwd), "field_backward": _time(field_bwd), "adam": _time(adam)}

    # the occupancy refresh runs every `update_interval` (16) steps inside the timed region of bench.py
    keep = (model.local_step, model.mean_count)
    phases["occupancy_refresh"] = _time(model.update_extra_state, reps=4, warm=1)
    phases["occupancy_refresh_per_step"] = phases["occupancy_refresh"] / max(trainer.update_interval, 1)
    model.local_step, model.mean_count = keep

    if F > 64 or C > 16:   # wide heads carve extra GEMM workspaces (csrc/field.cu::carve): phases only
        phases["live_samples"], phases["marched_samples"] = n_live, n_marched
        return {"roofline": None, "phases_ms": phases}
    # ---- candidate hot kernels, standalone, on the step's own buffers (layout: csrc/field.cu::carve)
    def carve_offsets():
        sizes = [M * desc.in_pad * 2, M * 16 * 4, M * 32 * 2, M * 16 * 2, M * (F + 16) * 2, M * (F + 16) * 4, M * F * 4,
                 M * 16 * 4, M * 16 * 4, M * 4 * 4, M * 16 * 4, int(desc.L) * M * 2 * 4, 16]
        offs, o = [], 0
        for s in sizes:
            offs.append(o)
            o += (s + 255) // 256 * 256
        return offs
    off = carve_offsets()
    base = fws.data_ptr()
    x_enc, h16, dout_sigma, d_enc, amax = base + off[0], base + off[1], base + off[10], base + off[11], base + off[12]
    hid = model.hidden_dim

    def k_encode():
        call("al_encode_position", ptr(p_xyzs if early else xyzs), M, ptr(p_meta if early else meta), float(model.bound), desc.encoding, desc.table, desc.offsets,
             desc.L, desc.S, desc.H, 0, ptr(p_xenc) if early else x_enc, desc.in_pad, st)

    def k_sigma_fwd():
        if early:
            call("al_mlp_forward", desc.in_pad, hid, 16, 2, desc.w_sigma, ptr(p_xenc), desc.in_pad, M, ptr(p_meta), ptr(p_h16),
                 16, 0, 0, 16, 0, ptr(p_sigma), 1, 0, 0, 1, 2, None, 0, 0, 0, 0, 0, st)
        else:
            call("al_mlp_forward", desc.in_pad, hid, 16, 2, desc.w_sigma, x_enc, desc.in_pad, M, ptr(meta), h16, 16, 0, 0, 16, 0,
                 ptr(vals), ldv, 0, 0, 1, 2, None, 0, 0, 0, 0, 0, st)

    def k_sigma_bwd():
        call("al_mlp_backward", desc.in_pad, hid, 16, 2, desc.w_sigma, x_enc, desc.in_pad, M, ptr(meta), dout_sigma, 16, 0, 16,
             ptr(amax_t), ptr(grads[1]), d_enc, 1, M, 12, 2 * int(desc.L), st)

    def k_scatter():
        call("al_grid_scatter_xyz", d_enc, M, ptr(xyzs), M, ptr(meta), float(model.bound), 1, desc.offsets, ptr(grads[0]),
             desc.L, desc.S, desc.H, 0, st)

    kern = {"encode_position": _time(k_encode), "sigma_mlp_forward": _time(k_sigma_fwd),
            "sigma_mlp_backward": _time(k_sigma_bwd), "grid_scatter": _time(k_scatter),
            "adam_table": _time(lambda: call("al_adam_step", ptr(shadow[0]), ptr(grads[0]), ptr(adam_state[0][0]),
                                             ptr(adam_state[0][1]), shadow[0].numel(), 5e-3, 0.9, 0.99, 1e-15, 0.0, 10,
                                             1.0, 1, st))}
    pk = peaks()
    in_pad, L = int(desc.in_pad), int(desc.L)
    mac_bwd = (in_pad * hid + hid * hid) + (16 * hid + hid * hid + hid * in_pad) + (in_pad * hid + hid * hid + hid * 16)
    mac_fwd = in_pad * hid + hid * hid + hid * 16
    work = {   # (bound, algorithmic units per launch, unit)
        "encode_position": ("hbm", n_marched * (12 + L * 8 * 8 + in_pad * 2) / 1e9, "GB/s"),
        "sigma_mlp_forward": ("tensor", n_marched * 2 * mac_fwd / 1e12, "TFLOP/s"),
        "sigma_mlp_backward": ("tensor", n_live * 2 * mac_bwd / 1e12, "TFLOP/s"),
        "grid_scatter": ("hbm", n_live * (12 + L * 8 + L * 8 * 8) / 1e9, "GB/s"),
        "adam_table": ("hbm", shadow[0].numel() * 32 / 1e9, "GB/s"),
    }
    top = max(kern, key=kern.get)
    bound, units, unit = work[top]
    achieved = units / (kern[top] * 1e-3)
    traffic, traffic_samples = committed_traffic(top) or (None, None)
    limiter = {"encode_position": "L1/L2 gather throughput, table L2-resident (ncu: l1tex and lts %, profiles/)",
               "grid_scatter": "L2 atomic throughput (ncu: lts %)", "adam_table": "HBM streams",
               "sigma_mlp_forward": "tensor pipe paced by TMEM reads", "sigma_mlp_backward": "tensor pipe paced by TMEM reads"}
    roofline = {"kernel": top, "bound": bound, "limiter": limiter.get(top), "achieved": achieved, "peak": pk[bound], "unit": unit,
                "frac": achieved / pk[bound], "traffic": traffic, "traffic_unit": "DRAM bytes per launch, ncu --set full capture of one step "
                "(profiles/roofline_traffic.json)", "traffic_marched_samples": traffic_samples, "peak_source": pk["source"],
                "avg_launch_ms": kern[top], "live_samples": n_live, "marched_samples": n_marched,
                "all": {k: {"ms": v, "bound": work[k][0], "achieved": work[k][1] / (v * 1e-3), "unit": work[k][2],
                            "frac": work[k][1] / (v * 1e-3) / pk[work[k][0]]} for k, v in kern.items()}}
    # the four heads' backward kernels as they run inside the step (k_mlp_bwd_tc2, two tiles in flight): field_backward
    # minus the hash-grid scatter, against the dense tensor peak on the MACs of recompute + dgrad + wgrad
    def bwd_macs(i, h, o, nh):
        hh = h * h if nh == 2 else 0
        return (i * h + hh) + (o * h + hh + h * i) + (i * h + hh + h * o)
    macs_all = bwd_macs(in_pad, hid, 16, 2) + bwd_macs(32, int(desc.hidden_color), 16, 2) + bwd_macs(16, F, F, 2) + bwd_macs(F + 16, 64, 16, 1)
    t_mlp_bwd = max(phases["field_backward"] - kern["grid_scatter"], 1e-6)
    tf = n_live * 2 * macs_all / (t_mlp_bwd * 1e-3) / 1e12
    roofline["all"]["mlp_backward_4_heads"] = {"ms": t_mlp_bwd, "bound": "tensor", "achieved": tf, "unit": "TFLOP/s", "frac": tf / pk["tensor"],
                                               "note": "field_backward - grid_scatter; MACs of recompute + dgrad + wgrad of the four heads"}
    if os.environ.get("AL_BWD_DBG_SWEEP"):
        # timing experiment (tools/job_bwd_dbg.sh): the same backward with pieces of the two-tile MLP kernels switched off
        sweep = {}
        for bits in (0, 1, 2, 3, 4, 5, 6, 7):
            lib.al_set_bwd_debug(bits)
            sweep[str(bits)] = _time(field_bwd) - kern["grid_scatter"]
        lib.al_set_bwd_debug(0)
        phases["mlp_backward_debug_sweep_ms"] = sweep
    if early:
        phases["density_pre"] = _time(density_pre)
        phases["compact_alive"] = _time(compact)
        phases["heads_forward"] = _time(heads_fwd)
    phases["live_samples"] = n_live
    phases["marched_samples"] = n_marched
    return {"roofline": roofline, "phases_ms": phases}


def measure_wide(args, scene, model, trainer, device, n_rays):
    """C5: the wide feature head (F = 512, csrc/gemm_tc.cu) standalone through the tcnn.Network operator on as many rows as a
    training step feeds it: achieved TFLOP/s of forward and forward+backward against the measured dense bf16 peak."""
    M = int(model.last_alive_meta[0].item()) if model.last_alive_meta is not None else n_rays * 256
    M = max(M // 128 * 128, 128)
    net = model.semantic_features
    x = torch.randn(M, net.n_input_dims, device=device) * 0.5
    with torch.no_grad():
        t_fwd = _time(lambda: net(x), reps=10, warm=2)
    xg = x.clone().requires_grad_(True)

    def fb():
        y = net(xg)
        y.backward(torch.ones_like(y) * 1e-3)
        net.params.grad = None
        xg.grad = None
    t_fb = _time(fb, reps=10, warm=2)
    macs = net.in_pad * net.hidden + (net.hidden * net.hidden if net.n_hidden == 2 else 0) + net.hidden * net.out_pad
    pk = peaks()
    fwd_tf = 2 * macs * M / (t_fwd * 1e-3) / 1e12
    fb_tf = 3 * 2 * macs * M / (t_fb * 1e-3) / 1e12
    return {"roofline": {"kernel": "k_gemm_tma (semantic_features %d->%d->%d->%d forward)" % (net.in_pad, net.hidden, net.hidden, net.out_pad),
                         "bound": "tensor", "achieved": fwd_tf, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": fwd_tf / pk["tensor"],
                         "traffic": None, "rows": M, "avg_launch_ms": t_fwd, "peak_source": pk["source"],
                         "forward_backward": {"ms": t_fb, "achieved": fb_tf, "frac": fb_tf / pk["tensor"]}}}
