"""numpy front-end of oracle/ngp_oracle.c (the CPU restatement of the reference's marching,
compositing and hash-grid index algorithms).

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg; never from autolabel_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ngp_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
SO = os.path.join(OUT_DIR, "libngp_oracle.so")

_lib = None


def build(force=False):
    """gcc -O2 -ffp-contract=off (no implicit FMA: the fused operations are spelled out)."""
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                               "-o", SO, SRC, "-lm"])
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


u32, f32 = C.c_uint32, C.c_float


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o, rays_d, aabb = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3), _f32(aabb)
    N = rays_o.shape[0]
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    ni, fi = np.empty(N, np.uint8), np.empty(N, np.uint8)
    lib().orc_near_far_from_aabb(_p(rays_o), _p(rays_d), _p(aabb), u32(N), f32(min_near), _p(nears), _p(fars),
                                 _p(ni), _p(fi))
    return nears, fars, ni, fi


def morton3D(coords):
    coords = _i32(coords)
    out = np.empty(coords.shape[0], np.int32)
    lib().orc_morton3D(_p(coords), u32(coords.shape[0]), _p(out))
    return out


def morton3D_invert(indices):
    indices = _i32(indices)
    out = np.empty((indices.shape[0], 3), np.int32)
    lib().orc_morton3D_invert(_p(indices), u32(indices.shape[0]), _p(out))
    return out


def packbits(grid, thresh):
    grid = _f32(grid).reshape(-1)
    N = grid.size // 8
    out = np.empty(N, np.uint8)
    lib().orc_packbits(_p(grid), u32(N), f32(thresh), _p(out))
    return out


def march_rays_train(rays_o, rays_d, bound, bitfield, C_, H, nears, fars, M, perturb=False, dt_gamma=0.0,
                     max_steps=1024):
    """Returns dict(xyzs, dirs, deltas, ts, rays, counter) with zero-initialised [M,...] buffers."""
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    nears, fars = _f32(nears), _f32(fars)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    N = rays_o.shape[0]
    xyzs, dirs = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32)
    deltas, ts = np.zeros((M, 2), np.float32), np.zeros((M,), np.float32)
    rays, counter = np.zeros((N, 3), np.int32), np.zeros(2, np.int32)
    lib().orc_march_rays_train(_p(rays_o), _p(rays_d), _p(bitfield), f32(bound), f32(dt_gamma), u32(max_steps),
                               u32(N), u32(C_), u32(H), u32(M), _p(nears), _p(fars), _p(xyzs), _p(dirs),
                               _p(deltas), _p(ts), _p(rays), _p(counter), u32(1 if perturb else 0))
    return dict(xyzs=xyzs, dirs=dirs, deltas=deltas, ts=ts, rays=rays, counter=counter)


def composite_rays_train_forward(sigmas, vals, deltas, rays, M=None):
    sigmas, vals, deltas, rays = _f32(sigmas), _f32(vals), _f32(deltas), _i32(rays)
    if vals.ndim == 1:
        vals = vals[:, None]
    K, N = vals.shape[1], rays.shape[0]
    M = sigmas.shape[0] if M is None else M
    ws, depth, image = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, K), np.float32)
    lib().orc_composite_rays_train_forward(_p(sigmas), _p(vals), u32(K), _p(deltas), _p(rays), u32(M), u32(N),
                                           _p(ws), _p(depth), _p(image))
    return ws, depth, image


def composite_rays_train_backward(g_ws, g_image, sigmas, vals, deltas, rays, ws, image, M=None):
    sigmas, vals, deltas, rays = _f32(sigmas), _f32(vals), _f32(deltas), _i32(rays)
    g_ws, g_image, ws, image = _f32(g_ws), _f32(g_image), _f32(ws), _f32(image)
    K, N = vals.shape[1], rays.shape[0]
    M = sigmas.shape[0] if M is None else M
    g_sig, g_vals = np.zeros_like(sigmas), np.zeros_like(vals)
    lib().orc_composite_rays_train_backward(_p(g_ws), _p(g_image), _p(sigmas), _p(vals), u32(K), _p(deltas),
                                            _p(rays), _p(ws), _p(image), u32(M), u32(N), _p(g_sig), _p(g_vals))
    return g_sig, g_vals


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C_, H, nears, fars,
               perturb=0, dt_gamma=0.0, max_steps=1024):
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    rays_alive, rays_t = _i32(rays_alive), _f32(rays_t)
    nears, fars = _f32(nears), _f32(fars)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    M = n_alive * n_step
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    lib().orc_march_rays(u32(n_alive), u32(n_step), _p(rays_alive), _p(rays_t), _p(rays_o), _p(rays_d), f32(bound),
                         f32(dt_gamma), u32(max_steps), u32(C_), u32(H), _p(bitfield), _p(nears), _p(fars),
                         _p(xyzs), _p(dirs), _p(deltas), u32(perturb))
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, vals, deltas, weights_sum, depth, image):
    """In place on rays_t, weights_sum, depth, image (float32 C-contiguous numpy arrays)."""
    sigmas, vals, deltas = _f32(sigmas), _f32(vals), _f32(deltas)
    K = vals.shape[1]
    lib().orc_composite_rays(u32(n_alive), u32(n_step), _p(_i32(rays_alive)), _p(rays_t), _p(sigmas), _p(vals),
                             u32(K), _p(deltas), _p(weights_sum), _p(depth), _p(image))


def compact_rays(n_alive, rays_alive_old, rays_t_old):
    rays_alive_old, rays_t_old = _i32(rays_alive_old), _f32(rays_t_old)
    rays_alive, rays_t = np.zeros_like(rays_alive_old), np.zeros_like(rays_t_old)
    counter = np.zeros(1, np.int32)
    lib().orc_compact_rays(u32(n_alive), _p(rays_alive), _p(rays_alive_old), _p(rays_t), _p(rays_t_old), _p(counter))
    return rays_alive, rays_t, int(counter[0])


def grid_offsets(num_levels=16, base_resolution=16, per_level_scale=2.0, log2_hashmap_size=19, input_dim=3):
    """Level offsets exactly as torch_ngp/gridencoder/grid.py:113-124."""
    offsets, offset = [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        params_in_level = min(max_params, (resolution + 1) ** input_dim)
        params_in_level = int(np.ceil(params_in_level / 8) * 8)
        offsets.append(offset)
        offset += params_in_level
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32)


def grid_encode_forward(inputs, table, offsets, per_level_scale, H, gridtype=0, level_scales=None):
    """outputs [L,B,C], indices [B,L,2^D] (gridencoder.cu:75-175); S = log2(per_level_scale) as grid.py:33."""
    inputs, table, offsets = _f32(inputs), _f32(table), _i32(offsets)
    B, D = inputs.shape
    Cc, L = table.shape[1], offsets.shape[0] - 1
    S = np.float32(np.log2(per_level_scale))
    out = np.zeros((L, B, Cc), np.float32)
    idx = np.zeros((B, L, 2 ** D), np.int32)
    ls = None if level_scales is None else _f32(level_scales)
    lib().orc_grid_encode_forward(_p(inputs), _p(table), _p(offsets), _p(out), u32(B), u32(D), u32(Cc), u32(L),
                                  f32(S), u32(H), u32(gridtype), _p(idx), _p(ls))
    return out, idx


def grid_encode_backward(grad, inputs, offsets, n_entries, per_level_scale, H, gridtype=0, level_scales=None):
    """grad [L,B,C] -> dense float64 table gradient [n_entries, C] (gridencoder.cu:226-312)."""
    grad, inputs, offsets = _f32(grad), _f32(inputs), _i32(offsets)
    L, B, Cc = grad.shape
    D = inputs.shape[1]
    S = np.float32(np.log2(per_level_scale))
    gg = np.zeros((n_entries, Cc), np.float64)
    ls = None if level_scales is None else _f32(level_scales)
    lib().orc_grid_encode_backward(_p(grad), _p(inputs), _p(offsets), _p(gg), u32(B), u32(D), u32(Cc), u32(L),
                                   f32(S), u32(H), u32(gridtype), _p(ls))
    return gg
