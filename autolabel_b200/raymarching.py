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
def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2, return_indices=False):
    """nears, fars [N] of the slab test against ``aabb`` (raymarching.py:21-60)."""
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


def march_rays_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None,
                     mean_count=-1, perturb=False, align=-1, force_all_rays=False, dt_gamma=0,
                     max_steps=1024):
    """Occupancy-guided sampling for training (raymarching.py:266-396).

    Returns ``xyzs [M,3], dirs [M,3], deltas [M,2], rays [N,3]`` with the reference's sizing rules:
    ``M = mean_count`` rounded up to ``align`` when known, else ``N * max_steps`` truncated to the
    counted total (one D2H read, as in the reference)."""
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
    r = _march_train_raw(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter, M, perturb,
                         dt_gamma, max_steps)
    xyzs, dirs, deltas = r["xyzs"], r["dirs"], r["deltas"]
    if force_all_rays or mean_count <= 0:
        m = int(r["counter"][0].item())
        if align > 0:
            m += align - m % align
        xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
    return xyzs, dirs, deltas, r["rays"]


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
