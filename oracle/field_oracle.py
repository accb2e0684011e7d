"""fp32 PyTorch restatement of the field (encodings + MLP heads) and of ragged compositing.

TEST INFRASTRUCTURE ONLY — tier O2 of DESIGN.md ("tolerance tier"): importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never from
autolabel_b200/.

PARITY UNPINNED for the tiny-cuda-nn parts: autolabel's MLPs and its Frequency / SphericalHarmonics /
Grid encodings come from tiny-cuda-nn (NVlabs), installed unpinned from git HEAD
(reference README.md:26) and absent from /root/reference and from this image.  The conventions
below restate tcnn's published behaviour and are anchored on the reference's call sites
(autolabel/models.py:15-59,84-136,150-256) and on the in-tree equivalents:
  * hash grid      -> torch_ngp/gridencoder/src/gridencoder.cu:35-223 with the hg+freq hyper-parameters
                      of models.py:39-48 (the compiled in-tree kernel pins indices, tests/golden)
  * SH degree 4    -> torch_ngp/shencoder/src/shencoder.cu:50-73, input mapped [0,1] -> [-1,1]
  * Frequency      -> sin/cos(2^k pi x), k < n, dimension-major, (sin, cos) interleaved
  * Network        -> bias-free Linear/ReLU stack, inputs padded with ONES to a multiple of 16 (tcnn pads
                      the encoded input with 1.0, so a padded column acts as a bias), outputs padded to 16
  * trunc_exp      -> torch_ngp/activation.py:1-17
The compositing restatement follows torch_ngp/nerf/renderer.py:243-311 on ragged (marched) segments.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

PRIMES = (1, 2654435761, 805459861)


def pad16(n):
    return (n + 15) // 16 * 16


# ------------------------------------------------------------------ encodings
def freq_encode(x, n_freq):
    """[B,D] -> [B, D*2*n_freq]: for each dim d, for each k: sin(2^k pi x_d), cos(2^k pi x_d)."""
    outs = []
    for d in range(x.shape[1]):
        for k in range(n_freq):
            a = x[:, d] * (2.0 ** k) * math.pi
            outs.append(torch.sin(a))
            outs.append(torch.cos(a))
    return torch.stack(outs, dim=1)


def sh4(d01):
    """tcnn SphericalHarmonics degree 4 of directions given in [0,1] (shencoder.cu:50-73)."""
    v = d01 * 2.0 - 1.0
    x, y, z = v[:, 0], v[:, 1], v[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    o = [
        torch.full_like(x, 0.28209479177387814),
        -0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x,
        1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.94617469575755997 * z2 - 0.31539156525251999,
        -1.0925484305920792 * xz, 0.54627421529603959 * x2 - 0.54627421529603959 * y2,
        0.59004358992664352 * y * (-3.0 * x2 + y2), 2.8906114426405538 * xy * z,
        0.45704579946446572 * y * (1.0 - 5.0 * z2), 0.3731763325901154 * z * (5.0 * z2 - 3.0),
        0.45704579946446572 * x * (1.0 - 5.0 * z2), 1.4453057213202769 * z * (x2 - y2),
        0.59004358992664352 * x * (-x2 + 3.0 * y2),
    ]
    return torch.stack(o, dim=1)


def grid_offsets(num_levels=16, base_resolution=16, per_level_scale=2.0, log2_hashmap_size=19, input_dim=3):
    """torch_ngp/gridencoder/grid.py:113-124."""
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


def grid_encode(x01, table, offsets, per_level_scale=2.0, H=16, gridtype=0):
    """Differentiable (w.r.t. table) restatement of gridencoder.cu:75-175 for D = 3.
    x01 [B,3] in [0,1] (out-of-range rows -> zeros), table [n,C]; returns [B, L*C] (level-major
    columns, the layout grid.py:51 returns)."""
    B = x01.shape[0]
    L = len(offsets) - 1
    C = table.shape[1]
    S = np.float32(np.log2(per_level_scale))
    oob = ((x01 < 0) | (x01 > 1)).any(dim=1)
    outs = []
    M32 = 0xFFFFFFFF
    for l in range(L):
        hs = int(offsets[l + 1] - offsets[l])
        scale = np.float32(np.exp2(np.float32(l) * S) * np.float32(H) - np.float32(1.0))
        res = int(math.ceil(float(scale))) + 1
        # fused multiply-add like the kernel (gridencoder.cu:134 compiles to FFMA): single rounding via float64
        pos = (x01.double() * float(scale) + 0.5).to(x01.dtype)
        pg = torch.floor(pos)
        fr = pos - pg
        pg = pg.to(torch.int64)
        acc = torch.zeros(B, C, dtype=table.dtype, device=table.device)
        for corner in range(8):
            w = torch.ones(B, dtype=x01.dtype, device=x01.device)
            q = []
            for d in range(3):
                if corner & (1 << d):
                    w = w * fr[:, d]
                    q.append(pg[:, d] + 1)
                else:
                    w = w * (1 - fr[:, d])
                    q.append(pg[:, d])
            stride, index, hashed = 1, torch.zeros(B, dtype=torch.int64, device=x01.device), False
            for d in range(3):
                if stride <= hs:
                    index = (index + q[d] * stride) & M32
                    stride = (stride * (res + 1)) & M32   # uint32 wrap-around, as in gridencoder.cu:56-63
            if gridtype == 0 and stride > hs:
                index = torch.zeros(B, dtype=torch.int64, device=x01.device)
                for d in range(3):
                    index = index ^ ((q[d] * PRIMES[d]) & M32)
            index = index % hs + int(offsets[l])
            index = torch.where(oob, torch.zeros_like(index), index)
            acc = acc + w[:, None] * table[index]
        acc = torch.where(oob[:, None], torch.zeros_like(acc), acc)
        outs.append(acc)
    return torch.cat(outs, dim=1)


# ------------------------------------------------------------------ MLP
def mlp_num_params(in_pad, hidden, out_pad, n_hidden):
    return hidden * in_pad + (hidden * hidden if n_hidden == 2 else 0) + out_pad * hidden


def mlp(x, params, in_pad, hidden, out_pad, n_hidden):
    """x [n, k<=in_pad] (ones-padded to in_pad), flat fp32 params [W1 | W2 | Wo], row-major [out, in]."""
    n = x.shape[0]
    if x.shape[1] < in_pad:
        x = torch.cat([x, torch.ones(n, in_pad - x.shape[1], dtype=x.dtype, device=x.device)], dim=1)
    o = 0
    W1 = params[o:o + hidden * in_pad].view(hidden, in_pad); o += hidden * in_pad
    h = F.relu(x @ W1.t())
    if n_hidden == 2:
        W2 = params[o:o + hidden * hidden].view(hidden, hidden); o += hidden * hidden
        h = F.relu(h @ W2.t())
    Wo = params[o:o + out_pad * hidden].view(out_pad, hidden)
    return h @ Wo.t()


class _TruncExp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _TruncExp.apply


# ------------------------------------------------------------------ field (ALNetwork, run() variant)
def encode_position(xyz, bound, encoding, table=None, offsets=None, per_level_scale=2.0, H=16):
    """autolabel/models.py:15-59,138-148.  encoding in {'freq', 'hg', 'hg+freq'}.
    (x + bound) / (2 bound) is evaluated as (x + bound) * (1 / (2 bound)), which is what torch does
    on CUDA for a division by a Python scalar."""
    inv = np.float32(1.0) / np.float32(2.0 * bound)
    xn = (xyz + bound) * float(inv)
    if encoding == 'freq':
        return freq_encode(xn, 10)
    if encoding == 'hg':
        return grid_encode(xn, table, offsets, per_level_scale, H)
    if encoding == 'hg+freq':
        return torch.cat([freq_encode(xyz, 2), grid_encode(xn.clamp(0.0, 1.0), table, offsets, per_level_scale, H)], dim=1)
    raise NotImplementedError(encoding)


def field_forward(xyz, dirs, P, cfg):
    """P: dict(table, w_sigma, w_color, w_semf, w_semo); cfg: dict(encoding, bound, hidden, hidden_color,
    feat_dim, n_classes, offsets, per_level_scale, H).  Returns sigma [n], rgb [n,3], logits [n,C],
    feat [n,F], h16 [n,16]  — density() / color() / semantic() of models.py:175-256."""
    Fd, Cc = cfg['feat_dim'], cfg['n_classes']
    x = encode_position(xyz, cfg['bound'], cfg['encoding'], P.get('table'), cfg.get('offsets'),
                        cfg.get('per_level_scale', 2.0), cfg.get('H', 16))
    in_pad = pad16(x.shape[1])
    h = mlp(x, P['w_sigma'], in_pad, cfg['hidden'], 16, 2)
    sigma = trunc_exp(h[:, 0])
    geo = h[:, 1:16]
    d01 = (dirs + 1) / 2
    hc = mlp(torch.cat([sh4(d01), geo], dim=1), P['w_color'], 32, cfg['hidden_color'], 16, 2)
    rgb = torch.sigmoid(hc[:, :3])
    feat = mlp(geo, P['w_semf'], 16, Fd, Fd, 2)
    logits = mlp(torch.cat([F.relu(feat), geo], dim=1), P['w_semo'], Fd + 16, 64, pad16(Cc), 1)[:, :Cc]
    return sigma, rgb, logits, feat, h


# ------------------------------------------------------------------ ragged compositing
def composite(sigmas, vals, deltas, tpos, xyzs, rays, M, sigma_scale=1.0):
    """renderer.py:243-311 on marched segments: alpha = 1 - exp(-sigma scale dt), T = prod(1 - alpha),
    w = alpha T;  returns weights_sum [N], depth = sum w t [N], depth_sq = sum w t^2, out = sum w vals
    [N,K], coords = sum w xyz [N,3].  Rays with count == 0 or offset + count >= M give zeros."""
    N = rays.shape[0]
    K = vals.shape[1]
    dev = sigmas.device
    ws, dep, dsq, out, crd = [], [], [], [], []
    rays_c = rays.cpu().numpy()
    order = np.argsort(rays_c[:, 0], kind='stable')
    res = {}
    for n in range(N):
        rid, off, cnt = int(rays_c[n, 0]), int(rays_c[n, 1]), int(rays_c[n, 2])
        if cnt == 0 or off + cnt >= M:
            z = torch.zeros((), device=dev, dtype=sigmas.dtype)
            res[rid] = (z, z, z, torch.zeros(K, device=dev, dtype=sigmas.dtype), torch.zeros(3, device=dev, dtype=sigmas.dtype))
            continue
        s = sigmas[off:off + cnt] * sigma_scale
        dt = deltas[off:off + cnt, 0]
        alpha = 1 - torch.exp(-s * dt)
        T = torch.cumprod(torch.cat([torch.ones(1, device=dev, dtype=alpha.dtype), 1 - alpha]), dim=0)[:-1]
        w = alpha * T
        t = tpos[off:off + cnt]
        res[rid] = (w.sum(), (w * t).sum(), (w * t * t).sum(), (w[:, None] * vals[off:off + cnt]).sum(0),
                    (w[:, None] * xyzs[off:off + cnt]).sum(0))
    del order
    ids = sorted(res)
    assert ids == list(range(N)), "ray ids must be a permutation of 0..N-1"
    return (torch.stack([res[i][0] for i in ids]), torch.stack([res[i][1] for i in ids]),
            torch.stack([res[i][2] for i in ids]), torch.stack([res[i][3] for i in ids]),
            torch.stack([res[i][4] for i in ids]))


def alive_prefix(sigmas, deltas, rays, M, sigma_scale=1.0, t_thresh=1e-4):
    """Training-time early termination (csrc/composite.cu::k_alive_count), restated with the rule of the reference's
    marched inference kernel (raymarching.cu:929-935: a ray stops after the sample that brings T below 1e-4): the
    number of leading samples of each ray whose transmittance BEFORE the sample is >= t_thresh, and that transmittance.
    Returns (counts [N] int64, T_before [M_total] float64).  Rays with count 0 or offset + count >= M stay empty."""
    rays_np = rays.detach().cpu().numpy()
    sig = sigmas.detach().double().cpu()
    dts = deltas.detach().double().cpu()[:, 0]
    counts = torch.zeros(rays_np.shape[0], dtype=torch.int64)
    T_before = torch.ones(sig.shape[0], dtype=torch.float64)
    for n, (rid, off, cnt) in enumerate(rays_np):
        if cnt == 0 or off + cnt >= M:
            continue
        alpha = 1 - torch.exp(-sig[off:off + cnt] * sigma_scale * dts[off:off + cnt])
        T = torch.cumprod(torch.cat([torch.ones(1, dtype=torch.float64), 1 - alpha]), 0)[:-1]
        T_before[off:off + cnt] = T
        dead = (T < t_thresh).nonzero()
        counts[n] = int(dead[0]) if dead.numel() else int(cnt)
    return counts, T_before


def render_outputs(ws, depth_raw, depth_sq, out, coords, direction_norms, n_classes, bg_color=1.0):
    """The per-ray epilogue of renderer.run() (renderer.py:270-320) on composited sums."""
    depth = depth_raw / direction_norms
    # sum w (depth - t)^2 = depth^2 sum w - 2 depth sum w t + sum w t^2   (depth already divided by the norm,
    # t not: exactly the mixed expression of renderer.py:277-278)
    depth_variance = (depth * depth * ws - 2 * depth * depth_raw + depth_sq).detach()
    image = out[:, :3] + (1 - ws)[:, None] * bg_color
    return {
        'depth': depth, 'depth_variance': depth_variance, 'image': image,
        'semantic': out[:, 3:3 + n_classes], 'semantic_features': out[:, 3 + n_classes:],
        'coordinates_map': coords,
    }


# ------------------------------------------------------------------ mixed-precision model of the kernel
def _q16(t):
    """Round to fp16 and back (the precision operands enter the tensor cores with)."""
    return t.half().float()


def mlp_fp16_model(x, params, in_pad, hidden, out_pad, n_hidden, dout=None, scale=1.0):
    """The arithmetic csrc/mlp.cu states it performs, in PyTorch: operands (inputs, weights, hidden
    activations, output gradients, hidden gradients) rounded to fp16, every product accumulated in
    fp32, ReLU masks taken from the fp16 activations.  Returns y and, if dout is given, (y, dx, dW)
    for dout scaled by `scale` on entry and unscaled on exit.  Used to separate 'the kernel implements
    its stated algorithm exactly' (tight check) from 'fp16 operands vs the fp32 oracle' (tolerance)."""
    n = x.shape[0]
    if x.shape[1] < in_pad:
        x = torch.cat([x, torch.ones(n, in_pad - x.shape[1], dtype=x.dtype, device=x.device)], dim=1)
    o = 0
    W1 = _q16(params[o:o + hidden * in_pad].view(hidden, in_pad)); o += hidden * in_pad
    W2 = None
    if n_hidden == 2:
        W2 = _q16(params[o:o + hidden * hidden].view(hidden, hidden)); o += hidden * hidden
    Wo = _q16(params[o:o + out_pad * hidden].view(out_pad, hidden))
    a0 = _q16(x).double()
    W1d, Wod = W1.double(), Wo.double()
    a1 = _q16(F.relu(a0 @ W1d.t()).float()).double()
    a_last = a1
    if n_hidden == 2:
        W2d = W2.double()
        a2 = _q16(F.relu(a1 @ W2d.t()).float()).double()
        a_last = a2
    y = (a_last @ Wod.t()).float()
    if dout is None:
        return y
    d = torch.zeros(n, out_pad, dtype=torch.float64, device=x.device)
    d[:, :dout.shape[1]] = _q16((dout * scale).clamp(-65504, 65504)).double()
    dl = _q16(((d @ Wod) * (a_last > 0)).float()).double()
    gWo = d.t() @ a_last
    if n_hidden == 2:
        d1 = _q16(((dl @ W2d) * (a1 > 0)).float()).double()
        gW2 = dl.t() @ a1
    else:
        d1 = dl
    gW1 = d1.t() @ a0
    dx = (d1 @ W1d) / scale
    parts = [gW1.reshape(-1)] + ([gW2.reshape(-1)] if n_hidden == 2 else []) + [gWo.reshape(-1)]
    dW = torch.cat(parts) / scale
    return y, dx.float(), dW.float()


def grad_scale_for(amax):
    """The power-of-two gradient scale csrc/mlp.cu derives on the device: largest 2^k with amax 2^k < 64."""
    if not (amax > 0) or not math.isfinite(amax):
        return 1.0
    m, e = math.frexp(amax)
    return float(2.0 ** max(-40, min(40, 6 - e)))
