"""Ray-marching operators — same public names, argument meaning and return values as the
reference's ``torch_ngp/raymarching/raymarching.py`` (:139-586), backed by the sm_100a kernels in
``csrc/raymarch.cu`` / ``csrc/composite.cu`` through the C ABI (``include/autolabel_b200.h``).

Differences a caller can observe (all documented in DESIGN.md):
  * segment offsets of ``march_rays_train`` are the exclusive scan of the per-ray counts in ray
    order (the reference's order depends on atomic scheduling, raymarching.cu:448-449);
  * ``composite_rays_train`` accepts any number of value channels (``rgbs`` may be ``[M, K]``)
    and its backward also propagates ``grad_depth`` (the reference drops it,
    raymarching.py:437-438);
  * kernels run on the current torch stream, not the legacy default stream.
"""
import torch
from torch.autograd import Function

from . import _lib
from ._lib import call, ptr, stream_ptr

__all__ = [
    "near_far_backward", "march_backward",
    "near_far_from_aabb", "polar_from_ray", "morton3D", "morton3D_invert", "packbits",
    "march_rays_train", "composite_rays_train", "composite_train_full", "march_rays",
    "composite_rays", "compact_rays",
]


def _f32c(t):
    """Geometry is always fp32, also under autocast (reference: custom_fwd(cast_inputs=float32))."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _cuda(t):
    return t if t.is_cuda else t.cuda()


# ------------------------------------------------------------------ utils
def _near_far_raw(rays_o, rays_d, aabb, min_near):
    rays_o = _f32c(_cuda(rays_o)).view(-1, 3)
    rays_d = _f32c(_cuda(rays_d)).view(-1, 3)
    aabb = _f32c(_cuda(aabb))
    N = rays_o.shape[0]
    nears = torch.empty(N, dtype=torch.float32, device=rays_o.device)
    fars = torch.empty(N, dtype=torch.float32, device=rays_o.device)
    ni = torch.empty(N, dtype=torch.uint8, device=rays_o.device)
    fi = torch.empty(N, dtype=torch.uint8, device=rays_o.device)
    call("al_near_far_from_aabb", ptr(rays_o), ptr(rays_d), ptr(aabb), N, float(min_near), ptr(nears),
         ptr(fars), ptr(ni), ptr(fi), stream_ptr(rays_o.device))
    return rays_o, rays_d, aabb, nears, fars, ni, fi


def near_far_backward(aabb, rays_o, rays_d, near_indices, far_indices, g_nears, g_fars):
    """Pose gradients of the slab test (raymarching.py:81-136): with the plane `aabb[idx]` on axis `idx % 3` that set
    the bound, t = (aabb[idx] - o[axis]) / d[axis], so dt/do[axis] = -1/d[axis] and dt/dd[axis] = (o[axis] -
    aabb[idx]) / d[axis]^2; rays that miss the box (index 255) get zero.  Like the reference, the clamp
    near = max(near, min_near) is not differentiated.  Pure torch: runs on any device."""
    N = rays_o.shape[0]
    rows = torch.arange(N, device=rays_o.device)
    g_o = torch.zeros_like(rays_o)
    g_d = torch.zeros_like(rays_d)
    for idx, g in ((near_indices, g_nears), (far_indices, g_fars)):
        if g is None:
            continue
        idx = idx.long()
        hit = (idx != 255).to(rays_o.dtype)
        axis = idx % 3
        plane = aabb[idx % 6]
        d_ax = rays_d[rows, axis]
        o_ax = rays_o[rows, axis]
        g = g.to(rays_o.dtype) * hit
        g_o[rows, axis] += g * (-1.0 / d_ax)
        g_d[rows, axis] += g * (o_ax - plane) / (d_ax * d_ax)
    return g_o, g_d


class _NearFar(Function):
    """near_far_from_aabb with the reference fork's backward w.r.t. the rays (raymarching.py:21-136)."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, aabb, min_near):
        rays_o, rays_d, aabb, nears, fars, ni, fi = _near_far_raw(rays_o, rays_d, aabb, min_near)
        ctx.save_for_backward(aabb, rays_o, rays_d, ni, fi)
        ctx.mark_non_differentiable(ni, fi)
        return nears, fars, ni, fi

    @staticmethod
    def backward(ctx, g_nears, g_fars, _gni, _gfi):
        aabb, rays_o, rays_d, ni, fi = ctx.saved_tensors
        g_o, g_d = near_far_backward(aabb, rays_o, rays_d, ni, fi, g_nears, g_fars)
        return g_o, g_d, None, None


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2, return_indices=False):
    """nears, fars [N] of the slab test against ``aabb`` (raymarching.py:21-60); differentiable w.r.t. the rays
    (pose gradients, raymarching.py:81-136) when they require grad."""
    if torch.is_grad_enabled() and (rays_o.requires_grad or rays_d.requires_grad):
        nears, fars, ni, fi = _NearFar.apply(rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), aabb, min_near)
    else:
        _, _, _, nears, fars, ni, fi = _near_far_raw(rays_o, rays_d, aabb, min_near)
    if return_indices:
        return nears, fars, ni, fi
    return nears, fars


def polar_from_ray(rays_o, rays_d, radius):
    raise NotImplementedError(
        "polar_from_ray (background sphere) is out of scope: autolabel asserts bg_radius <= 0 "
        "(torch_ngp/nerf/renderer.py:288-289)")


def morton3D(coords):
    """[N,3] int32 in [0,1024) -> [N] int32 Morton codes (raymarching.py:177-201)."""
    coords = _cuda(coords).int().contiguous()
    N = coords.shape[0]
    out = torch.empty(N, dtype=torch.int32, device=coords.device)
    call("al_morton3d", ptr(coords), N, ptr(out), stream_ptr(coords.device))
    return out


def morton3D_invert(indices):
    """[N] int32 -> [N,3] int32 (raymarching.py:204-227)."""
    indices = _cuda(indices).int().contiguous()
    N = indices.shape[0]
    out = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
    call("al_morton3d_invert", ptr(indices), N, ptr(out), stream_ptr(indices.device))
    return out


def packbits(grid, thresh, bitfield=None, thresh_dev=None):
    """bit i of byte n = grid.flat[8n+i] > thresh (raymarching.py:230-259).

    ``thresh_dev`` (optional 1-element CUDA tensor): the effective threshold becomes
    ``min(thresh, thresh_dev)`` without a host round trip."""
    grid = _f32c(_cuda(grid))
    N = grid.numel() // 8
    if bitfield is None:
        bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
    call("al_packbits", ptr(grid), N, float(thresh), ptr(thresh_dev), ptr(bitfield), stream_ptr(grid.device))
    return bitfield


# ------------------------------------------------------------------ training
def _march_train_raw(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter, M, perturb,
                     dt_gamma, max_steps, aabb=None, min_near=0.2, want_dirs=True, want_ts=True,
                     want_tpos=False, want_sray=False, zero_init=True):
    dev = rays_o.device
    N = rays_o.shape[0]
    alloc = torch.zeros if zero_init else torch.empty
    xyzs = alloc(M, 3, dtype=torch.float32, device=dev)
    dirs = alloc(M, 3, dtype=torch.float32, device=dev) if want_dirs else None
    deltas = alloc(M, 2, dtype=torch.float32, device=dev)
    ts = alloc(M, 1, dtype=torch.float32, device=dev) if want_ts else None
    tpos = alloc(M, dtype=torch.float32, device=dev) if want_tpos else None
    sray = alloc(M, dtype=torch.int32, device=dev) if want_sray else None
    rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
    meta = torch.zeros(2, dtype=torch.int32, device=dev)
    if step_counter is None:
        step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
    ws_bytes = _lib.lib.al_march_rays_train_workspace(N, int(max_steps))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    nears_out = fars_out = None
    if nears is None:
        nears_out = torch.empty(N, dtype=torch.float32, device=dev)
        fars_out = torch.empty(N, dtype=torch.float32, device=dev)
    call("al_march_rays_train", ptr(rays_o), ptr(rays_d), ptr(density_bitfield), float(bound), float(dt_gamma),
         int(max_steps), N, int(C), int(H), int(M), ptr(nears), ptr(fars), ptr(aabb), float(min_near),
         ptr(nears_out), ptr(fars_out), ptr(xyzs), ptr(dirs), ptr(deltas), ptr(ts), ptr(tpos), ptr(sray),
         ptr(rays), ptr(step_counter), ptr(meta), 1 if perturb else 0, ptr(ws), stream_ptr(dev))
    return dict(xyzs=xyzs, dirs=dirs, deltas=deltas, ts=ts, tpos=tpos, sray=sray, rays=rays, meta=meta,
                nears=nears if nears is not None else nears_out, fars=fars if fars is not None else fars_out,
                counter=step_counter)


def march_backward(rays, ts, g_xyzs, g_dirs, n_rays):
    """Pose gradients of march_rays_train (raymarching.py:358-392): every sample is x = o + t d and carries the ray's
    direction, so dL/do = sum over the ray's samples of dL/dx and dL/dd = sum of (dL/dx * ts + dL/ddirs), with `ts`
    the operator's own ts output, as in the reference.  Segments follow `rays` = (ray id, offset, count); samples
    outside every segment (padding, dropped rays) contribute nothing.  Pure torch: runs on any device."""
    counts = rays[:, 2].long()
    offsets = rays[:, 1].long()
    total = int(g_xyzs.shape[0]) if g_xyzs is not None else int(g_dirs.shape[0])
    keep = (counts > 0) & (offsets + counts <= total)
    counts = torch.where(keep, counts, torch.zeros_like(counts))
    n_live = int(counts.sum().item())
    ray_of = torch.repeat_interleave(torch.arange(rays.shape[0], device=rays.device), counts, output_size=n_live)
    first = torch.repeat_interleave(offsets, counts, output_size=n_live)
    start_in_list = torch.repeat_interleave(torch.cumsum(counts, 0) - counts, counts, output_size=n_live)
    sample = first + (torch.arange(n_live, device=rays.device) - start_in_list)
    ids = rays[:, 0].long()[ray_of]
    dt = g_xyzs.dtype if g_xyzs is not None else g_dirs.dtype
    g_o = torch.zeros(n_rays, 3, dtype=dt, device=rays.device)
    g_d = torch.zeros(n_rays, 3, dtype=dt, device=rays.device)
    if g_xyzs is not None:
        gx = g_xyzs[sample]
        g_o.index_add_(0, ids, gx)
        g_d.index_add_(0, ids, gx * ts.reshape(-1, 1)[sample].to(dt))
    if g_dirs is not None:
        g_d.index_add_(0, ids, g_dirs[sample])
    return g_o, g_d


class _MarchRaysTrain(Function):
    """march_rays_train with the reference fork's backward w.r.t. the rays (raymarching.py:266-396)."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter, M, perturb, dt_gamma,
                max_steps):
        r = _march_train_raw(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter, M, perturb,
                             dt_gamma, max_steps)
        ctx.save_for_backward(r["rays"], r["ts"])
        ctx.n_rays = rays_o.shape[0]
        ctx.mark_non_differentiable(r["rays"], r["counter"])
        return r["xyzs"], r["dirs"], r["deltas"], r["rays"], r["counter"]

    @staticmethod
    def backward(ctx, g_xyzs, g_dirs, _g_deltas, _g_rays, _g_counter):
        rays, ts = ctx.saved_tensors
        g_o, g_d = march_backward(rays, ts, g_xyzs, g_dirs, ctx.n_rays)
        return (g_o, g_d) + (None,) * 11


def march_rays_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None,
                     mean_count=-1, perturb=False, align=-1, force_all_rays=False, dt_gamma=0,
                     max_steps=1024):
    """Occupancy-guided sampling for training (raymarching.py:266-396).

    Returns ``xyzs [M,3], dirs [M,3], deltas [M,2], rays [N,3]`` with the reference's sizing rules:
    ``M = mean_count`` rounded up to ``align`` when known, else ``N * max_steps`` truncated to the
    counted total (one D2H read, as in the reference).  Differentiable w.r.t. the rays (pose gradients,
    raymarching.py:358-392) when they require grad."""
    needs_grad = torch.is_grad_enabled() and (rays_o.requires_grad or rays_d.requires_grad)
    rays_o = _f32c(_cuda(rays_o)).view(-1, 3)
    rays_d = _f32c(_cuda(rays_d)).view(-1, 3)
    density_bitfield = _cuda(density_bitfield).contiguous()
    nears = _f32c(_cuda(nears))
    fars = _f32c(_cuda(fars))
    N = rays_o.shape[0]
    M = N * max_steps
    if not force_all_rays and mean_count > 0:
        if align > 0:
            mean_count += align - mean_count % align
        M = mean_count
    if needs_grad:
        xyzs, dirs, deltas, rays, counter = _MarchRaysTrain.apply(rays_o, rays_d, bound, density_bitfield, C, H,
                                                                  nears.detach(), fars.detach(), step_counter, M, perturb,
                                                                  dt_gamma, max_steps)
    else:
        r = _march_train_raw(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter, M, perturb,
                             dt_gamma, max_steps)
        xyzs, dirs, deltas, rays, counter = r["xyzs"], r["dirs"], r["deltas"], r["rays"], r["counter"]
    if force_all_rays or mean_count <= 0:
        m = int(counter[0].item())
        if align > 0:
            m += align - m % align
        xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
    return xyzs, dirs, deltas, rays


class _CompositeTrain(Function):
    """K-channel compositing with full backward (sigma, values; depth and opacity gradients)."""

    @staticmethod
    def forward(ctx, sigmas, vals, deltas, rays, tpos, xyzs, sigma_scale, M_cap):
        sigmas = _f32c(sigmas)
        vals = _f32c(vals)
        if vals.dim() == 1:
            vals = vals.view(-1, 1)
        deltas = _f32c(deltas)
        M = sigmas.shape[0] if M_cap is None else int(M_cap)
        N = rays.shape[0]
        K = vals.shape[1]
        dev = sigmas.device
        ws = torch.empty(N, dtype=torch.float32, device=dev)
        depth = torch.empty(N, dtype=torch.float32, device=dev)
        depth_sq = torch.empty(N, dtype=torch.float32, device=dev)
        out = torch.empty(N, K, dtype=torch.float32, device=dev)
        coords = torch.empty(N, 3, dtype=torch.float32, device=dev) if xyzs is not None else None
        call("al_composite_train_fwd", ptr(sigmas), sigmas.stride(0), ptr(vals), vals.stride(0), K, ptr(deltas),
             ptr(tpos), ptr(xyzs), ptr(rays), M, N, float(sigma_scale), ptr(ws), ptr(depth), ptr(depth_sq),
             ptr(out), ptr(coords), stream_ptr(dev))
        ctx.save_for_backward(sigmas, vals, deltas, rays, ws, depth, out, tpos if tpos is not None else torch.empty(0))
        ctx.cfg = (M, N, K, float(sigma_scale), tpos is not None)
        ctx.mark_non_differentiable(depth_sq)
        if coords is None:
            coords = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(coords)
        return ws, depth, depth_sq, out, coords

    @staticmethod
    def backward(ctx, g_ws, g_depth, _g_sq, g_out, _g_coords):
        sigmas, vals, deltas, rays, ws, depth, out, tpos = ctx.saved_tensors
        M, N, K, scale, has_t = ctx.cfg
        dev = sigmas.device
        g_out = torch.zeros_like(out) if g_out is None else _f32c(g_out)
        g_ws = None if g_ws is None else _f32c(g_ws)
        g_depth = None if g_depth is None else _f32c(g_depth)
        g_sigmas = torch.zeros_like(sigmas)
        g_vals = torch.zeros_like(vals)
        call("al_composite_train_bwd", ptr(g_ws), ptr(g_depth), ptr(g_out), ptr(sigmas), sigmas.stride(0),
             ptr(vals), vals.stride(0), K, ptr(deltas), ptr(tpos) if has_t else None, ptr(rays), ptr(ws),
             ptr(depth), ptr(out), M, N, scale, ptr(g_sigmas), g_sigmas.stride(0), ptr(g_vals),
             g_vals.stride(0), None, stream_ptr(dev))
        return g_sigmas, g_vals, None, None, None, None, None, None


def composite_train_full(sigmas, vals, deltas, rays, tpos=None, xyzs=None, sigma_scale=1.0, M=None):
    """(weights_sum [N], depth [N], depth_sq [N], out [N,K], coords [N,3] or empty)."""
    return _CompositeTrain.apply(sigmas, vals, deltas, rays, tpos, xyzs, sigma_scale, M)


def composite_rays_train(sigmas, rgbs, deltas, rays):
    """weights_sum [N], depth [N], image [N,K] (raymarching.py:399-457)."""
    ws, depth, _sq, image, _c = _CompositeTrain.apply(sigmas, rgbs, deltas, rays, None, None, 1.0, None)
    return ws, depth, image


# ------------------------------------------------------------------ inference
def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far,
               align=-1, perturb=False, dt_gamma=0, max_steps=1024, return_extra=False):
    """March ``n_step`` samples for each alive ray (raymarching.py:464-533)."""
    rays_o = _f32c(_cuda(rays_o)).view(-1, 3)
    rays_d = _f32c(_cuda(rays_d)).view(-1, 3)
    dev = rays_o.device
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    xyzs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
    dirs = torch.zeros(M, 3, dtype=torch.float32, device=dev)
    deltas = torch.zeros(M, 2, dtype=torch.float32, device=dev)
    tpos = torch.zeros(M, dtype=torch.float32, device=dev) if return_extra else None
    call("al_march_rays", int(n_alive), int(n_step), ptr(rays_alive), ptr(rays_t), ptr(rays_o), ptr(rays_d),
         float(bound), float(dt_gamma), int(max_steps), int(C), int(H), ptr(density_bitfield), ptr(near),
         ptr(far), ptr(xyzs), ptr(dirs), ptr(deltas), ptr(tpos), None, int(perturb), stream_ptr(dev))
    if return_extra:
        return xyzs, dirs, deltas, tpos
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image,
                   tpos=None, xyzs=None, depth_sq=None, coords=None, sigma_scale=1.0):
    """In-place accumulation into ``weights_sum / depth / image`` (raymarching.py:536-562);
    ``image`` may carry any number of channels ``[N, K]`` with ``rgbs [n_alive*n_step, K]``."""
    sigmas = _f32c(sigmas)
    rgbs = _f32c(rgbs)
    K = rgbs.shape[1]
    call("al_composite_rays", int(n_alive), int(n_step), ptr(rays_alive), ptr(rays_t), ptr(sigmas),
         sigmas.stride(0), ptr(rgbs), rgbs.stride(0), K, ptr(deltas), ptr(tpos), ptr(xyzs), float(sigma_scale),
         ptr(weights_sum), ptr(depth), ptr(depth_sq), ptr(image), ptr(coords), stream_ptr(sigmas.device))
    return tuple()


def compact_rays(n_alive, rays_alive, rays_alive_old, rays_t, rays_t_old, alive_counter):
    """Remove dead rays (rays_t_old < 0), keeping order (raymarching.py:565-586)."""
    call("al_compact_rays", int(n_alive), ptr(rays_alive), ptr(rays_alive_old), ptr(rays_t), ptr(rays_t_old),
         ptr(alive_counter), stream_ptr(rays_alive.device))
    return tuple()
