"""Volume renderer — same class interface as the reference's ``torch_ngp/nerf/renderer.py``
(``NeRFRenderer`` :68-744): buffers, ``run`` / ``run_cuda`` / ``render`` /
``mark_untrained_grid`` / ``update_extra_state`` / ``reset_extra_state``.

``run_cuda`` is the B200 path (disabled by two ``assert(False)`` in the reference, :331,:697): fused
slab test + occupancy-guided marching -> fused field (encoding + MLP heads) -> K-channel
compositing, returning the SAME dictionary as ``run()`` (:313-320): depth (metric: sum w t / |d|),
depth_variance, image (white background), semantic logits, semantic_features, coordinates_map.
Training keeps the reference's sample-budget contract (``mean_count`` rounded up to 128, rays
whose segment overflows are dropped, raymarching.cu:459) and is free of host synchronisation.
"""
import math
import os

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib, raymarching
from ._lib import FieldDesc, call, ptr, stream_ptr

import ctypes


def _nonzero_known(mask, count):
    """torch.nonzero(mask).squeeze(-1) when the number of hits is already known on the host: no device synchronisation."""
    try:
        return torch.nonzero_static(mask, size=count).squeeze(-1)
    except (RuntimeError, NotImplementedError, AttributeError):
        return torch.nonzero(mask).squeeze(-1)


def _round_up(v, a):
    return (v + a - 1) // a * a


class _TrainCtx:
    """Buffers of one marched training forward, kept for the backward (fused path)."""
    __slots__ = ("M", "N", "K", "ldv", "xyzs", "deltas", "tpos", "sray", "rays", "meta", "vals", "fws", "ws", "depth",
                 "depth_sq", "out", "coords", "training")


class StepArena:
    """Named scratch buffers that outlive a step: `get` returns a view of a cached flat buffer and only allocates when
    the request outgrows it.  The graph-captured training step records kernels on these addresses, so a re-capture for a
    new sample budget allocates nothing (allocation during stream capture is refused)."""

    def __init__(self, device):
        self.device = device
        self.bufs = {}

    def get(self, name, shape, dtype=torch.float32):
        if isinstance(shape, int):
            shape = (shape,)
        numel = 1
        for d in shape:
            numel *= int(d)
        buf = self.bufs.get(name)
        if buf is None or buf.dtype != dtype or buf.numel() < numel:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError(f"StepArena: buffer '{name}' would have to grow during graph capture")
            buf = torch.empty(max(numel, 1), dtype=dtype, device=self.device)
            self.bufs[name] = buf
        return buf[:numel].view(*shape)


class _TorchAlloc:
    """StepArena interface on plain torch allocations (eager steps)."""

    def __init__(self, device):
        self.device = device

    def get(self, name, shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=self.device)


def fused_train_forward(model, rays_o, rays_d, M, perturb, dt_gamma, max_steps, counter, training, arena=None,
                        budget_dev=None):
    """march -> field -> composite on one stream, no host synchronisation (every kernel reads the live sample
    count from device memory).  Returns a _TrainCtx holding the per-ray outputs and what the backward needs.
    `budget_dev` (int32 [1] on the device, optional): M is then only the CAPACITY of the sample buffers and the
    reference's overflow rule (raymarching.cu:458-459) is applied with min(M, budget_dev[0]) inside the march."""
    dev = rays_o.device
    N = rays_o.shape[0]
    st = stream_ptr(dev)
    K = model.n_channels
    ldv = 1 + K
    desc = model.field_desc()
    A = arena if arena is not None else _TorchAlloc(dev)
    c = _TrainCtx()
    c.M, c.N, c.K, c.ldv, c.training = M, N, K, ldv, training
    thr = float(getattr(model, 'train_t_thresh', 0.0))
    early = thr > 0.0
    pre = 'pre_' if early else ''          # with early termination the march fills staging buffers
    xyzs = A.get(pre + 'xyzs', (M, 3))
    deltas = A.get(pre + 'deltas', (M, 2))
    tpos = A.get(pre + 'tpos', M)
    sray = A.get(pre + 'sray', M, torch.int32)
    rays = A.get(pre + 'rays', (N, 3), torch.int32)
    meta = A.get(pre + 'meta', 2, torch.int32)
    mws = A.get('march_ws', _lib.lib.al_march_rays_train_workspace(N, max_steps), torch.uint8)
    aabb = model.aabb_train if model.training else model.aabb_infer
    call("al_march_rays_train_budget", ptr(rays_o), ptr(rays_d), ptr(model.density_bitfield), float(model.bound),
         float(dt_gamma), int(max_steps), N, int(model.cascade), int(model.grid_size), int(M), ptr(budget_dev), None,
         None, ptr(aabb), float(model.min_near), None, None, ptr(xyzs), None, ptr(deltas), None, ptr(tpos),
         ptr(sray), ptr(rays), ptr(counter), ptr(meta), 1 if perturb else 0, ptr(mws), st)
    del mws
    c.vals = A.get('vals', (M, ldv))
    c.fws = A.get('field_ws', _lib.lib.al_field_workspace(ctypes.byref(desc), M, 1 if training else 0), torch.uint8)
    model.last_meta = meta                 # {samples written, samples counted} of the march
    if not early:
        c.xyzs, c.deltas, c.tpos, c.sray, c.rays, c.meta = xyzs, deltas, tpos, sray, rays, meta
        call("al_field_forward", ctypes.byref(desc), ptr(c.xyzs), ptr(rays_d), ptr(c.sray), M, ptr(c.meta), ptr(c.vals),
             ldv, None, 0, ptr(c.fws), st)
    else:
        # training-time early termination: density on every marched sample, then only the alive prefix of each ray
        # (transmittance before the sample >= train_t_thresh) goes through the heads, the compositing and the backward
        in_pad = int(desc.in_pad)
        x_enc = A.get('pre_x_enc', (M, in_pad), torch.float16)
        h16 = A.get('pre_h16', (M, 16))
        sigma = A.get('pre_sigma', M)
        call("al_field_density_pre", ctypes.byref(desc), ptr(xyzs), M, ptr(meta), ptr(x_enc), ptr(h16), ptr(sigma), st)
        c.xyzs = A.get('xyzs', (M, 3))
        c.deltas = A.get('deltas', (M, 2))
        c.tpos = A.get('tpos', M)
        c.sray = A.get('sray', M, torch.int32)
        c.rays = A.get('rays', (N, 3), torch.int32)
        c.meta = A.get('meta', 2, torch.int32)
        alive_ws = A.get('alive_ws', N, torch.int32)
        slot_x, slot_h = ctypes.c_void_p(), ctypes.c_void_p()
        call("al_field_workspace_slots", ctypes.byref(desc), M, 1 if training else 0, ptr(c.fws), ctypes.byref(slot_x),
             ctypes.byref(slot_h))
        call("al_compact_alive", ptr(sigma), ptr(deltas), ptr(rays), M, N, float(model.density_scale), thr, ptr(xyzs),
             ptr(tpos), ptr(sray), ptr(x_enc), in_pad, ptr(h16), ptr(c.rays), ptr(c.meta), ptr(c.xyzs), ptr(c.deltas),
             ptr(c.tpos), ptr(c.sray), slot_x, slot_h, ptr(c.vals), ldv, ptr(alive_ws), st)
        call("al_field_heads_forward", ctypes.byref(desc), ptr(rays_d), ptr(c.sray), M, ptr(c.meta), ptr(c.vals), ldv,
             ptr(c.fws), st)
    model.last_alive_meta = c.meta
    c.ws = A.get('ws', N)
    c.depth = A.get('depth', N)
    c.depth_sq = A.get('depth_sq', N)
    c.out = A.get('out', (N, K))
    c.coords = A.get('coords', (N, 3))
    call("al_composite_train_fwd", ptr(c.vals), ldv, c.vals.data_ptr() + 4, ldv, K, ptr(c.deltas), ptr(c.tpos),
         ptr(c.xyzs), ptr(c.rays), M, N, float(model.density_scale), ptr(c.ws), ptr(c.depth), ptr(c.depth_sq),
         ptr(c.out), ptr(c.coords), st)
    return c


def fused_train_backward(model, c, g_ws, g_depth, g_out, params, arena=None):
    """Backward of fused_train_forward: parameter gradients are accumulated by the kernels directly into
    ``param.grad`` (allocated here when missing): no 57 MB temporary per step for the hash table, and the gradient
    buffer doubles as the all-reduce buffer."""
    M, N, K, ldv = c.M, c.N, c.K, c.ldv
    dev = c.xyzs.device
    st = stream_ptr(dev)
    desc = model.field_desc()
    A = arena if arena is not None else _TorchAlloc(dev)
    amax = A.get('amax', 1)
    amax.zero_()
    grads = []
    for p in params:
        if p is not None and p.requires_grad:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
            grads.append(p.grad)
        else:
            grads.append(None)
    g_table, g_sigma, g_color, g_semf, g_semo = grads
    vals = c.vals
    if _lib.lib.al_set_mlp_backend(-1) == 1:
        # rank-1 backward: dL/dvals[i, c] = w[i] * g_out[ray(i), c] is never materialised (8 B / sample instead
        # of 4 (1 + K)); the tcgen05 head kernels rebuild their output gradients on the fly.
        w_s = A.get('w_samples', M)
        g_sig = A.get('g_sigma_samples', M)
        call("al_composite_train_bwd_weights", ptr(g_ws), ptr(g_depth), ptr(g_out), ptr(vals), ldv,
             vals.data_ptr() + 4, ldv, K, ptr(c.deltas), ptr(c.tpos), ptr(c.rays), ptr(c.ws), ptr(c.depth), ptr(c.out),
             M, N, float(model.density_scale), ptr(w_s), ptr(g_sig), ptr(amax), st)
        call("al_field_backward_rays", ctypes.byref(desc), ptr(c.xyzs), M, ptr(c.meta), ptr(vals), ldv, ptr(w_s),
             ptr(g_sig), ptr(g_out), ptr(c.sray), ptr(amax), ptr(g_table), ptr(g_sigma), ptr(g_color), ptr(g_semf),
             ptr(g_semo), ptr(c.fws), st)
    else:
        g_vals = A.get('g_vals', (M, ldv))
        call("al_composite_train_bwd", ptr(g_ws), ptr(g_depth), ptr(g_out), ptr(vals), ldv, vals.data_ptr() + 4,
             ldv, K, ptr(c.deltas), ptr(c.tpos), ptr(c.rays), ptr(c.ws), ptr(c.depth), ptr(c.out), M, N,
             float(model.density_scale), ptr(g_vals), ldv, g_vals.data_ptr() + 4, ldv, ptr(amax), st)
        call("al_field_backward", ctypes.byref(desc), ptr(c.xyzs), M, ptr(c.meta), ptr(vals), ptr(g_vals), ptr(amax),
             ldv, ptr(g_table), ptr(g_sigma), ptr(g_color), ptr(g_semf), ptr(g_semo), ptr(c.fws), st)


class _FusedRender(Function):
    """march -> field -> composite as ONE autograd node (fused_train_forward / fused_train_backward); ``None`` is
    returned for the parameter inputs because their gradients are accumulated in place."""

    @staticmethod
    def forward(ctx, model, rays_o, rays_d, M, perturb, dt_gamma, max_steps, counter, *params):
        training = any(p is not None and p.requires_grad for p in params)
        c = fused_train_forward(model, rays_o, rays_d, M, perturb, dt_gamma, max_steps, counter, training)
        if training:
            ctx.model, ctx.c, ctx.params = model, c, params
        ctx.mark_non_differentiable(c.depth_sq, c.coords)
        return c.ws, c.depth, c.depth_sq, c.out, c.coords

    @staticmethod
    def backward(ctx, g_ws, g_depth, _g_sq, g_out, _g_coords):
        c = ctx.c
        dev = c.xyzs.device
        g_out = torch.zeros(c.N, c.K, dtype=torch.float32, device=dev) if g_out is None else g_out.float().contiguous()
        g_ws = None if g_ws is None else g_ws.float().contiguous()
        g_depth = None if g_depth is None else g_depth.float().contiguous()
        fused_train_backward(ctx.model, c, g_ws, g_depth, g_out, ctx.params)
        ctx.c = None
        return (None,) * 8 + (None,) * len(ctx.params)


class NeRFRenderer(nn.Module):

    def __init__(self, bound=1, cuda_ray=False, density_scale=1, min_near=0.2, density_thresh=0.01,
                 bg_radius=-1):
        super().__init__()
        if hasattr(bound, 'shape'):
            bound = float(np.abs(bound[1] - bound[0]).max())
        self.bound = bound
        self.cascade = 1 + math.ceil(math.log2(self.bound))
        self.grid_size = 128
        self.density_scale = density_scale
        self.min_near = min_near
        self.density_thresh = density_thresh
        self.bg_radius = bg_radius
        if bg_radius > 0:
            raise NotImplementedError("bg_radius > 0 is not supported (the reference asserts it off, renderer.py:288-289)")
        aabb = torch.tensor([-bound, -bound, -bound, bound, bound, bound], dtype=torch.float32)
        self.register_buffer('aabb_train', aabb)
        self.register_buffer('aabb_infer', aabb.clone())
        self.cuda_ray = cuda_ray
        if cuda_ray:
            self.register_buffer('density_grid', torch.zeros([self.cascade, self.grid_size ** 3]))
            self.register_buffer('density_bitfield',
                                 torch.zeros(self.cascade * self.grid_size ** 3 // 8, dtype=torch.uint8))
            self.mean_density = 0
            self.iter_density = 0
            self.register_buffer('step_counter', torch.zeros(16, 2, dtype=torch.int32))
            self.mean_count = 0
            self.local_step = 0
        self.last_meta = None
        self.last_alive_meta = None
        # training: 0 (default) composites EVERY marched sample, exactly as the reference's training kernels do
        # (raymarching.cu:547-740; their `if (weight < 1e-4f) break` is commented out, :594,:696).  Opt-in
        # (opt.train_t_thresh / bench.py --train-t-thresh): a ray's samples behind the point where its transmittance
        # drops below this value skip the heads / compositing / backward (the rule of the inference kernel, :929-935).
        self.train_t_thresh = 0.0
        self.max_render_rays = 1 << 19  # rays per fused inference pass (bounds scratch memory)
        self.early_termination = True   # inference: stop a ray once its transmittance drops below 1e-4
        # samples marched per alive ray in successive waves: short waves while most rays are alive (a ray's samples behind
        # its termination point are wasted field evaluations, half a wave per ray on average), long ones for the few rays
        # that keep travelling through empty space
        self.wave_steps = (32, 32, 32, 32, 32, 64, 128, 256, 512)    # swept on the bench scene (profiles/r4r_wave_sweep.md)
        if os.environ.get('AL_WAVE_STEPS'):                # tuning aid: comma-separated samples per wave
            self.wave_steps = tuple(int(v) for v in os.environ['AL_WAVE_STEPS'].split(','))
        self.max_wave_samples = 1 << 24
        # inference waves: compositing folded into the head kernels' epilogues when the heads are the weight-resident
        # tcgen05 shapes (al_field_heads_forward_sum); AL_FUSED_COMPOSITE=0 keeps the value-matrix path
        self.fused_composite = os.environ.get('AL_FUSED_COMPOSITE', '1') != '0'
        self.max_scratch_bytes = 24 << 30   # per-pass scratch budget of the inference paths (vals + field workspace)

    # ------------------------------------------------------------ hooks implemented by the model
    def forward(self, x, d):
        raise NotImplementedError()

    def density(self, x):
        raise NotImplementedError()

    def color(self, x, d, mask=None, **kwargs):
        raise NotImplementedError()

    def reset_extra_state(self):
        if not self.cuda_ray:
            return
        self.density_grid.zero_()
        self.mean_density = 0
        self.iter_density = 0
        self.step_counter.zero_()
        self.mean_count = 0
        self.local_step = 0

    # ------------------------------------------------------------ reference (PyTorch) sampling path
    def run(self, rays_o, rays_d, direction_norms, num_steps=256, upsample_steps=0, bg_color=None,
            perturb=False, **kwargs):
        """Uniform sampling + PyTorch compositing, the path the reference executes today
        (renderer.py:186-320); density / color / semantic run on the sm_100a modules."""
        assert upsample_steps == 0
        assert bg_color is None
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        direction_norms = direction_norms.contiguous().view(-1)
        N = rays_o.shape[0]
        dev = rays_o.device
        aabb = self.aabb_train if self.training else self.aabb_infer
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, aabb, self.min_near)
        nears = nears.unsqueeze(-1)
        fars = fars.unsqueeze(-1)
        z = torch.linspace(0.0, 1.0, num_steps, device=dev).unsqueeze(0).expand((N, num_steps))
        z = nears + (fars - nears) * z
        sample_dist = (fars - nears) / num_steps
        if perturb:
            z = z + (torch.rand(z.shape, device=dev) - 0.5) * sample_dist
        xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * z.unsqueeze(-1)
        xyzs = torch.min(torch.max(xyzs, aabb[:3]), aabb[3:])
        dens = self.density(xyzs.reshape(-1, 3))
        sigma = dens['sigma'].view(N, num_steps)
        geo = dens['geo_feat'].view(N, num_steps, -1)
        deltas = torch.cat([z[..., 1:] - z[..., :-1], sample_dist * torch.ones_like(z[..., :1])], dim=-1)
        alphas = 1 - torch.exp(-deltas * self.density_scale * sigma)
        shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-15], dim=-1)
        weights = alphas * torch.cumprod(shifted, dim=-1)[..., :-1]
        dirs = rays_d.view(-1, 1, 3).expand_as(xyzs)
        mask = weights > 1e-4
        rgbs = self.color(xyzs.reshape(-1, 3), dirs.reshape(-1, 3), mask=mask.reshape(-1),
                          geo_feat=geo.reshape(-1, geo.shape[-1]), sigma=sigma.reshape(-1, 1)).view(N, -1, 3)
        weights = weights * mask
        weights_sum = weights.sum(dim=-1)
        depth = (weights * z).sum(dim=-1) / direction_norms
        depth_variance = (weights * (depth[..., None] - z) ** 2).sum(dim=-1).detach()
        w = weights.unsqueeze(-1)
        coordinates_map = (w * xyzs).sum(dim=-2)
        image = (w * rgbs).sum(dim=-2) + (1 - weights_sum).unsqueeze(-1) * 1
        semantic, sem_feat = self.semantic(geo.reshape(-1, geo.shape[-1]), sigma.reshape(-1, 1))
        semantic = (w * semantic.view(N, num_steps, self.semantic_classes)).sum(dim=-2)
        sem_feat = (w * sem_feat.view(N, num_steps, -1)).sum(dim=-2)
        return {
            'depth': depth.view(*prefix), 'depth_variance': depth_variance, 'image': image.view(*prefix, 3),
            'semantic': semantic, 'semantic_features': sem_feat, 'coordinates_map': coordinates_map,
        }

    # ------------------------------------------------------------ B200 path
    def _epilogue(self, ws, depth_raw, depth_sq, out, coords, direction_norms, bg_color, prefix):
        C = self.semantic_classes
        norms = direction_norms.reshape(-1).to(ws.dtype)
        depth = depth_raw / norms
        depth_variance = (depth * depth * ws - 2 * depth * depth_raw + depth_sq).detach()
        if bg_color is None:
            bg_color = 1
        image = out[:, :3] + (1 - ws).unsqueeze(-1) * bg_color
        return {
            'depth': depth.view(*prefix), 'depth_variance': depth_variance, 'image': image.view(*prefix, 3),
            'semantic': out[:, 3:3 + C], 'semantic_features': out[:, 3 + C:], 'coordinates_map': coords,
        }

    def run_cuda(self, rays_o, rays_d, direction_norms=None, dt_gamma=0, bg_color=None, perturb=False,
                 force_all_rays=False, max_steps=1024, **kwargs):
        if not self.cuda_ray:
            raise RuntimeError("run_cuda needs cuda_ray=True")
        if not rays_o.is_cuda:
            raise RuntimeError("run_cuda needs CUDA tensors; there is no CPU fallback")
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3).float()
        rays_d = rays_d.contiguous().view(-1, 3).float()
        N = rays_o.shape[0]
        if direction_norms is None:
            direction_norms = torch.ones(N, device=rays_o.device)
        params = self.field_params()

        if self.training:
            counter = self.step_counter[self.local_step % 16]
            counter.zero_()
            self.local_step += 1
            M = N * max_steps
            if not force_all_rays and self.mean_count > 0:
                # raymarching.py:324-327: mean_count += align - mean_count % align (align = 128)
                M = self.mean_count + 128 - self.mean_count % 128
            res = _FusedRender.apply(self, rays_o, rays_d, M, perturb, dt_gamma, max_steps, counter, *params)
            return self._epilogue(*res, direction_norms, bg_color, prefix)

        # inference: waves of samples with early termination (renderer.py:403-472); `early_termination=False`
        # composites every occupied sample instead (exact sample budget per chunk, one D2H read per chunk)
        outs = []
        waves = kwargs.get('early_termination', self.early_termination)
        with torch.no_grad():
            for head in range(0, N, self.max_render_rays):
                ro = rays_o[head:head + self.max_render_rays]
                rd = rays_d[head:head + self.max_render_rays]
                if waves:
                    outs.append(self._render_waves(ro, rd, perturb, dt_gamma, max_steps))
                else:
                    outs.append(self._render_chunk(ro, rd, perturb, dt_gamma, max_steps))
        res = [torch.cat([o[i] for o in outs], dim=0) for i in range(5)]
        return self._epilogue(*res, direction_norms, bg_color, prefix)

    def sample_budget(self, n_rays, max_steps=1024, force_all_rays=False):
        """raymarching.py:324-327: `mean_count` rounded up to the next multiple of 128 once known, else N * max_steps."""
        if not force_all_rays and self.mean_count > 0:
            return self.mean_count + 128 - self.mean_count % 128
        return n_rays * max_steps

    def budget_tensor(self, M):
        """The sample budget as a device scalar (int32 [1], fixed address): rewritten only when the value changes."""
        dev = self.density_bitfield.device
        if getattr(self, '_budget_dev', None) is None or self._budget_dev.device != dev:
            self._budget_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            self._budget_host = None
        if self._budget_host != int(M):
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("the sample budget must be written before the step is captured, not inside the graph")
            self._budget_dev.fill_(int(M))
            self._budget_host = int(M)
        return self._budget_dev

    def train_forward_raw(self, rays_o, rays_d, dt_gamma=0, perturb=True, force_all_rays=False, max_steps=1024,
                          counter=None, arena=None, capacity=None):
        """The marched training forward without autograd (SimpleTrainer's fused step): same sample-budget rule
        as run_cuda.  Returns the _TrainCtx; `fused_train_backward` consumes it.  `counter` (int32 [2], zeroed by the
        caller) replaces the rotating step-counter row: the graph-captured step copies it back after the replay."""
        rays_o = rays_o.contiguous().view(-1, 3).float()
        rays_d = rays_d.contiguous().view(-1, 3).float()
        N = rays_o.shape[0]
        if counter is None:
            counter = self.step_counter[self.local_step % 16]
            counter.zero_()
            self.local_step += 1
        M = self.sample_budget(N, max_steps, force_all_rays)
        budget = None
        if capacity is not None and capacity >= M:
            # buffers (and the captured launch geometry) sized `capacity`, the budget itself ALWAYS read from device
            # memory: a replayed graph follows later changes of mean_count (the reference's overflow rule,
            # raymarching.cu:458-459) also when capacity == M at capture time
            budget, M = self.budget_tensor(M), int(capacity)
        return fused_train_forward(self, rays_o, rays_d, M, perturb, dt_gamma, max_steps, counter, True, arena,
                                   budget_dev=budget), rays_d

    def _sample_bytes(self, desc):
        """Scratch bytes per marched sample in inference: vals row + field workspace + sample record."""
        return 4 * (1 + self.n_channels) + _lib.lib.al_field_workspace(ctypes.byref(desc), 4096, 0) // 4096 + 40

    def _wave_buffers(self, dev, N, cap, ldv, desc):
        """Sample buffers of the wave loop, allocated once per (rays, capacity) and reused by every wave and frame
        (the reference allocates and zero-fills them per iteration, raymarching.py:520-524; here the march marks an
        exhausted ray itself and `sray` only ever holds valid ray ids, so nothing is cleared between waves)."""
        key = (str(dev), int(N), int(cap), int(ldv))
        wb = getattr(self, '_wave_cache', None)
        if wb is None or wb['key'] != key:
            f32 = dict(dtype=torch.float32, device=dev)
            wb = {'key': key, 'xyzs': torch.zeros(cap, 3, **f32), 'deltas': torch.zeros(cap, 2, **f32),
                  'tpos': torch.zeros(cap, **f32), 'sray': torch.zeros(cap, dtype=torch.int32, device=dev),
                  'vals': torch.empty(cap, ldv, **f32), 'sigma': torch.empty(cap, **f32), 'w': torch.empty(cap, **f32),
                  'fws': torch.empty(_lib.lib.al_field_workspace(ctypes.byref(desc), cap, 0), dtype=torch.uint8, device=dev),
                  'alive': torch.empty(2, N, dtype=torch.int32, device=dev), 'rays_t': torch.empty(2, N, **f32),
                  'nears': torch.empty(N, **f32), 'fars': torch.empty(N, **f32),
                  'counter': torch.zeros(1, dtype=torch.int32, device=dev),
                  'counter_host': torch.empty(1, dtype=torch.int32, pin_memory=True),
                  'arange': torch.arange(N, dtype=torch.int32, device=dev)}
            self._wave_cache = wb
        return wb

    def fused_wave_composite(self):
        """True when the inference waves can fold compositing into the head epilogues: tcgen05 back end and the
        weight-resident head shapes (csrc/field.cu: feat_dim 64, at most 16 classes)."""
        return (self.fused_composite and _lib.lib.al_set_mlp_backend(-1) == 1 and self.hidden_dim_semantic == 64
                and self.semantic_classes <= 16)

    def _render_waves(self, rays_o, rays_d, perturb, dt_gamma, max_steps):
        """The reference's inference loop (renderer.py:403-472: march_rays -> field -> composite_rays ->
        compact_rays until no ray is alive) with the field fused and a wave schedule of 32 samples per alive ray
        (64 .. 512 once few rays are left) instead of 1-8 samples per iteration: a few large launches on preallocated
        buffers and one 4-byte D2H read (the alive count) per wave.  A ray stops after the sample that starts with
        transmittance < 1e-4 (raymarching.cu:929-935); what a wave marches beyond that point is discarded, so short
        early waves keep the wasted field evaluations near half a wave per ray."""
        dev = rays_o.device
        st = stream_ptr(dev)
        N = rays_o.shape[0]
        K = self.n_channels
        ldv = 1 + K
        desc = self.field_desc()
        f32 = dict(dtype=torch.float32, device=dev)
        aabb = self.aabb_train if self.training else self.aabb_infer
        wave_cap = max(1, min(self.max_wave_samples, self.max_scratch_bytes // self._sample_bytes(desc)))
        cap = int(min(wave_cap, N * self.wave_steps[0]))
        wb = self._wave_buffers(dev, N, cap, ldv, desc)
        nears, fars, alive, rays_t, counter = wb['nears'], wb['fars'], wb['alive'], wb['rays_t'], wb['counter']
        xyzs, deltas, tpos, sray, vals, fws = wb['xyzs'], wb['deltas'], wb['tpos'], wb['sray'], wb['vals'], wb['fws']
        call("al_near_far_from_aabb", ptr(rays_o), ptr(rays_d), ptr(aabb), N, float(self.min_near), ptr(nears),
             ptr(fars), None, None, st)
        ws, depth, depth_sq = torch.zeros(N, **f32), torch.zeros(N, **f32), torch.zeros(N, **f32)
        out, coords = torch.zeros(N, K, **f32), torch.zeros(N, 3, **f32)
        alive[0].copy_(wb['arange'])
        rays_t[0].copy_(nears)
        n_alive, step, i, total = N, 0, 0, 0
        fused = self.fused_wave_composite()
        while step < max_steps:
            cur, nxt = i % 2, (i + 1) % 2
            if i > 0:
                counter.zero_()
                call("al_compact_rays", n_alive, ptr(alive[cur]), ptr(alive[nxt]), ptr(rays_t[cur]), ptr(rays_t[nxt]),
                     ptr(counter), st)
                wb['counter_host'].copy_(counter, non_blocking=True)
                torch.cuda.current_stream(dev).synchronize()
                n_alive = int(wb['counter_host'][0])
            if n_alive <= 0:
                break
            n_step = self.wave_steps[min(i, len(self.wave_steps) - 1)]
            n_step = max(1, min(n_step, max_steps - step, cap // n_alive))
            M = n_alive * n_step
            call("al_march_rays", n_alive, n_step, ptr(alive[cur]), ptr(rays_t[cur]), ptr(rays_o), ptr(rays_d),
                 float(self.bound), float(dt_gamma), int(max_steps), int(self.cascade), int(self.grid_size),
                 ptr(self.density_bitfield), ptr(nears), ptr(fars), ptr(xyzs), None, ptr(deltas), ptr(tpos), ptr(sray),
                 1 if perturb else 0, st)
            if fused:
                # density -> weights (ray-level sums, stopping rule) -> heads that add w * value straight into `out`
                call("al_field_density_inputs", ctypes.byref(desc), ptr(xyzs), ptr(rays_d), ptr(sray), M, None,
                     ptr(wb['sigma']), ptr(fws), st)
                call("al_composite_rays_weights", n_alive, n_step, ptr(alive[cur]), ptr(rays_t[cur]), ptr(wb['sigma']), 1,
                     ptr(deltas), ptr(tpos), ptr(xyzs), float(self.density_scale), ptr(ws), ptr(depth), ptr(depth_sq),
                     ptr(coords), ptr(wb['w']), st)
                call("al_field_heads_forward_sum", ctypes.byref(desc), ptr(rays_d), ptr(sray), M, None, ptr(wb['w']),
                     ptr(out), K, 1, ptr(fws), st)
            else:
                call("al_field_forward", ctypes.byref(desc), ptr(xyzs), ptr(rays_d), ptr(sray), M, None, ptr(vals), ldv,
                     None, 0, ptr(fws), st)
                call("al_composite_rays", n_alive, n_step, ptr(alive[cur]), ptr(rays_t[cur]), ptr(vals), ldv,
                     vals.data_ptr() + 4, ldv, K, ptr(deltas), ptr(tpos), ptr(xyzs), float(self.density_scale), ptr(ws),
                     ptr(depth), ptr(depth_sq), ptr(out), ptr(coords), st)
            total += M
            step += n_step
            i += 1
        self.last_meta = torch.tensor([total, total], dtype=torch.int32)
        return ws, depth, depth_sq, out, coords

    def _render_chunk(self, rays_o, rays_d, perturb, dt_gamma, max_steps):
        dev = rays_o.device
        st = stream_ptr(dev)
        N = rays_o.shape[0]
        K = self.n_channels
        ldv = 1 + K
        desc = self.field_desc()
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)
        meta = torch.empty(2, dtype=torch.int32, device=dev)
        mws = torch.empty(_lib.lib.al_march_rays_train_workspace(N, max_steps), dtype=torch.uint8, device=dev)
        aabb = self.aabb_train if self.training else self.aabb_infer
        big = 0xFFFFFFFF
        call("al_march_rays_train_count", ptr(rays_o), ptr(rays_d), ptr(self.density_bitfield), float(self.bound),
             float(dt_gamma), int(max_steps), N, int(self.cascade), int(self.grid_size), big, None, None, ptr(aabb),
             float(self.min_near), None, None, ptr(rays), None, ptr(meta), 1 if perturb else 0, ptr(mws), st)
        total = int(meta[1].item())
        if N > 1 and total * self._sample_bytes(desc) > self.max_scratch_bytes:
            # too many samples for one pass (wide feature heads): halve the ray chunk
            del mws, rays
            h = N // 2
            a = self._render_chunk(rays_o[:h], rays_d[:h], perturb, dt_gamma, max_steps)
            b = self._render_chunk(rays_o[h:], rays_d[h:], perturb, dt_gamma, max_steps)
            return tuple(torch.cat([x, y], dim=0) for x, y in zip(a, b))
        M = _round_up(total + 1, 128)
        xyzs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        deltas = torch.empty(M, 2, dtype=torch.float32, device=dev)
        tpos = torch.empty(M, dtype=torch.float32, device=dev)
        sray = torch.empty(M, dtype=torch.int32, device=dev)
        call("al_march_rays_train_write", ptr(rays_o), ptr(rays_d), float(self.bound), float(dt_gamma),
             int(max_steps), N, int(self.cascade), int(self.grid_size), M, ptr(rays), ptr(xyzs), None, ptr(deltas),
             None, ptr(tpos), ptr(sray), ptr(mws), st)
        vals = torch.empty(M, ldv, dtype=torch.float32, device=dev)
        fws = torch.empty(_lib.lib.al_field_workspace(ctypes.byref(desc), M, 0), dtype=torch.uint8, device=dev)
        n_live = torch.full((1,), total, dtype=torch.int32, device=dev)
        call("al_field_forward", ctypes.byref(desc), ptr(xyzs), ptr(rays_d), ptr(sray), M, ptr(n_live), ptr(vals),
             ldv, None, 0, ptr(fws), st)
        ws = torch.empty(N, dtype=torch.float32, device=dev)
        depth = torch.empty(N, dtype=torch.float32, device=dev)
        depth_sq = torch.empty(N, dtype=torch.float32, device=dev)
        out = torch.empty(N, K, dtype=torch.float32, device=dev)
        coords = torch.empty(N, 3, dtype=torch.float32, device=dev)
        call("al_composite_train_fwd", ptr(vals), ldv, vals.data_ptr() + 4, ldv, K, ptr(deltas), ptr(tpos),
             ptr(xyzs), ptr(rays), M, N, float(self.density_scale), ptr(ws), ptr(depth), ptr(depth_sq), ptr(out),
             ptr(coords), st)
        self.last_meta = meta
        return ws, depth, depth_sq, out, coords

    # ------------------------------------------------------------ occupancy grid
    @torch.no_grad()
    def mark_untrained_grid(self, poses, intrinsic, S=64):
        """Cells never seen by any camera are set to -1 (renderer.py:479-561): one kernel over (cascade, cell)
        instead of the reference's five nested Python loops."""
        if not self.cuda_ray:
            return
        if isinstance(poses, np.ndarray):
            poses = torch.from_numpy(poses)
        dev = self.density_grid.device
        if dev.type != 'cuda':
            raise RuntimeError("mark_untrained_grid needs the model on a CUDA device; there is no CPU fallback")
        poses = poses.to(dev).float().contiguous()
        fx, fy, cx, cy = [float(v) for v in intrinsic]
        call("al_mark_untrained_grid", ptr(self.density_grid), ptr(poses), int(poses.shape[0]), fx, fy, cx, cy,
             float(self.bound), int(self.cascade), int(self.grid_size), stream_ptr(dev))

    @torch.no_grad()
    def mark_untrained_grid_torch(self, poses, intrinsic, S=64):
        """The reference's formulation (renderer.py:479-561) in torch ops; test oracle for the kernel above."""
        if isinstance(poses, np.ndarray):
            poses = torch.from_numpy(poses)
        dev = self.density_grid.device
        poses = poses.to(dev).float()
        B = poses.shape[0]
        fx, fy, cx, cy = [float(v) for v in intrinsic]
        H = self.grid_size
        count = torch.zeros_like(self.density_grid)
        ar = torch.arange(H, dtype=torch.int32, device=dev)
        for xs in ar.split(S):
            for ys in ar.split(S):
                for zs in ar.split(S):
                    xx, yy, zz = torch.meshgrid(xs, ys, zs, indexing='ij')
                    coords = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1), zz.reshape(-1, 1)], dim=-1)
                    indices = raymarching.morton3D(coords).long()
                    world = (2 * coords.float() / (H - 1) - 1).unsqueeze(0)
                    for cas in range(self.cascade):
                        bound = min(2 ** cas, self.bound)
                        half = bound / H
                        cas_world = world * (bound - half)
                        head = 0
                        while head < B:
                            tail = min(head + S, B)
                            cam = cas_world - poses[head:tail, :3, 3].unsqueeze(1)
                            cam = cam @ poses[head:tail, :3, :3]
                            mz = cam[:, :, 2] > 0
                            mx = torch.abs(cam[:, :, 0]) < cx / fx * cam[:, :, 2] + half * 2
                            my = torch.abs(cam[:, :, 1]) < cy / fy * cam[:, :, 2] + half * 2
                            count[cas, indices] += (mz & mx & my).sum(0).reshape(-1)
                            head += S
        return count == 0

    @torch.no_grad()
    def update_extra_state(self, decay=0.95, S=128):
        """Occupancy refresh (renderer.py:563-683): jittered density queries in Morton order,
        grid = max(grid * decay, new) on valid cells, mean, bit-packing, mean sample count.
        The sampling positions use the same torch RNG calls, shapes and order as the reference, so
        identical seeds give identical query points; the density query is the fused encoder +
        density MLP; EMA-max + mean + threshold + packbits run without leaving the device."""
        if not self.cuda_ray:
            return
        dev = self.density_grid.device
        H = self.grid_size
        # ONE host read-back per refresh: the sum of the step counters (-> mean_count, renderer.py:677-680) and, after the
        # first 16 refreshes, the number of occupied cells per cascade (the `high` of the reference's randint and the
        # size of its nonzero(), renderer.py:630-637).  density_grid does not change until the end of this function,
        # so reading the counts up front gives the values the reference reads cascade by cascade.
        total_step = min(16, self.local_step)
        want = []
        if total_step > 0:
            want.append(self.step_counter[:total_step, 0].sum(dtype=torch.int64).view(1))
        if self.iter_density >= 16:
            want.append((self.density_grid > 0).sum(dim=1, dtype=torch.int64))
        host = torch.cat(want).tolist() if want else []
        count_sum = host.pop(0) if total_step > 0 else 0
        occ_counts = host
        tmp_grid = -torch.ones_like(self.density_grid)
        if self.iter_density < 16:
            ar = torch.arange(H, dtype=torch.int32, device=dev)
            for xs in ar.split(S):
                for ys in ar.split(S):
                    for zs in ar.split(S):
                        xx, yy, zz = torch.meshgrid(xs, ys, zs, indexing='ij')
                        coords = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1), zz.reshape(-1, 1)], dim=-1)
                        indices = raymarching.morton3D(coords).long()
                        xyzs = 2 * coords.float() / (H - 1) - 1
                        for cas in range(self.cascade):
                            bound = min(2 ** cas, self.bound)
                            half = bound / H
                            cas_xyzs = xyzs * (bound - half)
                            cas_xyzs += (torch.rand_like(cas_xyzs) * 2 - 1) * half
                            sigmas = self.density_only(cas_xyzs)
                            sigmas *= self.density_scale
                            tmp_grid[cas, indices] = sigmas
        else:
            N = H ** 3 // 4
            for cas in range(self.cascade):
                coords = torch.randint(0, H, (N, 3), device=dev)
                indices = raymarching.morton3D(coords).long()
                occ_indices = _nonzero_known(self.density_grid[cas] > 0, int(occ_counts[cas]))
                rand_mask = torch.randint(0, occ_indices.shape[0], [N], dtype=torch.long, device=dev)
                occ_indices = occ_indices[rand_mask]
                occ_coords = raymarching.morton3D_invert(occ_indices)
                indices = torch.cat([indices, occ_indices], dim=0)
                coords = torch.cat([coords, occ_coords], dim=0)
                xyzs = 2 * coords.float() / (H - 1) - 1
                bound = min(2 ** cas, self.bound)
                half = bound / H
                cas_xyzs = xyzs * (bound - half)
                cas_xyzs += (torch.rand_like(cas_xyzs) * 2 - 1) * half
                sigmas = self.density_only(cas_xyzs)
                sigmas *= self.density_scale
                tmp_grid[cas, indices] = sigmas

        mean = torch.empty(1, dtype=torch.float32, device=dev)
        call("al_density_grid_update", ptr(self.density_grid), ptr(tmp_grid), self.density_grid.numel(),
             float(decay), ptr(mean), stream_ptr(dev))
        self.mean_density_dev = mean
        self.iter_density += 1
        raymarching.packbits(self.density_grid, self.density_thresh, self.density_bitfield, thresh_dev=mean)

        if total_step > 0:
            self.mean_count = int(count_sum / total_step)
        self.local_step = 0

    @property
    def mean_density(self):
        dev_val = getattr(self, 'mean_density_dev', None)
        if dev_val is not None:
            return float(dev_val.item())
        return self._mean_density

    @mean_density.setter
    def mean_density(self, v):
        self._mean_density = v
        self.mean_density_dev = None

    # ------------------------------------------------------------ staging (renderer.py:685-744)
    def render(self, rays_o, rays_d, direction_norms, staged=False, max_ray_batch=4096, **kwargs):
        if self.cuda_ray:
            # the marched path is never staged by the caller (renderer.py:704); it chunks internally
            return self.run_cuda(rays_o, rays_d, direction_norms, **kwargs)
        _run = self.run
        B, N = rays_o.shape[:2]
        dev = rays_o.device
        if not staged:
            return _run(rays_o, rays_d, direction_norms, **kwargs)
        res = {
            'depth': torch.empty((B, N), device=dev),
            'depth_variance': torch.empty((B, N), device=dev),
            'image': torch.empty((B, N, 3), device=dev),
            'semantic': torch.empty((B, N, self.semantic_classes), device=dev),
            'semantic_features': torch.empty((B, N, self.hidden_dim_semantic), device=dev),
            'coordinates_map': torch.empty((B, N, 3), device=dev),
        }
        for b in range(B):
            head = 0
            while head < N:
                tail = min(head + max_ray_batch, N)
                r = _run(rays_o[b:b + 1, head:tail], rays_d[b:b + 1, head:tail],
                         direction_norms[b:b + 1, head:tail], **kwargs)
                for k in res:
                    res[k][b:b + 1, head:tail] = r[k]
                head += max_ray_batch
        return res
