"""Synthetic RGB-D scenes of the shapes BASELINE.json names: an analytic indoor
room (axis-aligned box with textured walls + spheres), inward-facing orbit cameras, RGB / metric
depth / sparse semantic labels / feature maps rendered analytically.  Bench / test infrastructure
(no datasets can be downloaded here); everything is generated from seeds on the chosen device.

The sampler mirrors the batch contract of the reference's dataset (autolabel/dataset.py:182-242):
a batch is `batch_size // 512` chunks of 512 rays, each chunk from one image, pixels drawn with
replacement, sub-pixel jitter on the ray direction, and the dict keys
{rays_o, rays_d, direction_norms, pixels, depth, semantic, features}.
"""
import math

import numpy as np
import torch

ROOM = 1.5            # half extent of the room (metres): extents 3 m -> bound 3.0, 3 cascades
SPHERES = [((0.36, -0.96, 0.18), 0.48), ((-0.6, -1.14, -0.54), 0.36), ((0.12, -1.2, -0.84), 0.3)]


def _intersect(o, d):
    """Analytic ray cast. o, d [n,3] (d unit). Returns t [n], normal [n,3], obj id [n] (0 walls, 1.. spheres)."""
    n = o.shape[0]
    dev = o.device
    inv = 1.0 / torch.where(d.abs() < 1e-9, torch.full_like(d, 1e-9), d)
    t1 = (-ROOM - o) * inv
    t2 = (ROOM - o) * inv
    tfar = torch.maximum(t1, t2)                  # exit distance per axis (camera is inside the room)
    t_wall, axis = tfar.min(dim=1)
    normal = torch.zeros(n, 3, device=dev)
    normal.scatter_(1, axis[:, None], -torch.sign(d.gather(1, axis[:, None])))
    t = t_wall
    obj = torch.zeros(n, dtype=torch.long, device=dev)
    for i, (c, r) in enumerate(SPHERES):
        c = torch.tensor(c, device=dev)
        oc = o - c
        b = (oc * d).sum(1)
        cc = (oc * oc).sum(1) - r * r
        disc = b * b - cc
        ts = -b - torch.sqrt(disc.clamp(min=0))
        hit = (disc > 0) & (ts > 1e-3) & (ts < t)
        t = torch.where(hit, ts, t)
        nrm = (o + ts[:, None] * d - c) / r
        normal = torch.where(hit[:, None], nrm, normal)
        obj = torch.where(hit, torch.full_like(obj, i + 1), obj)
    return t, normal, obj


def _shade(p, normal, obj):
    """Procedural albedo + lambert term in [0,1]."""
    tex = 0.5 + 0.5 * torch.sin(p * 3.1) * torch.cos(p.roll(1, dims=1) * 2.3)
    base = torch.stack([0.35 + 0.15 * obj.float(), 0.55 - 0.1 * obj.float(), 0.4 + 0.1 * (obj % 2).float()], dim=1)
    light = torch.tensor([0.3, 0.8, 0.5], device=p.device)
    light = light / light.norm()
    lam = 0.55 + 0.45 * (normal * light).sum(1, keepdim=True).clamp(min=0)
    return (0.6 * base + 0.4 * tex).clamp(0, 1) * lam


class SyntheticScene:
    """n_frames images of height x width; all arrays live on `device`.

        images    [n, H*W, 3] fp32      depths [n, H*W] fp32 (metres, z-depth; 0 = invalid)
        semantics [n, H*W] int64 (-1 = unlabeled, classes 0..n_classes-1, `label_frac` of the pixels)
        features  [n, fh*fw, F] fp16    (feature maps at reduced resolution, dino: 90x120 for 480x640)
        poses     [n, 4, 4] camera-to-world, intrinsics (fx, fy, cx, cy)
    """

    def __init__(self, n_frames=300, height=480, width=640, feature_dim=64, feature_hw=None, n_classes=2,
                 label_frac=0.05, seed=0, device='cuda', chunk=512, lazy=False):
        self.n, self.h, self.w, self.F, self.C = n_frames, height, width, feature_dim, n_classes
        self.device = torch.device(device)
        self.chunk = chunk
        self.fx = self.fy = 0.8 * width
        self.cx, self.cy = width / 2.0, height / 2.0
        self.intrinsics = (self.fx, self.fy, self.cx, self.cy)
        self.min_bounds = np.array([-ROOM] * 3, dtype=np.float32)
        self.max_bounds = np.array([ROOM] * 3, dtype=np.float32)
        g = torch.Generator().manual_seed(seed)
        self.gen = torch.Generator(device=self.device).manual_seed(seed + 1)
        self.poses = self._make_poses(g).to(self.device)
        self.proj = (torch.randn(8, feature_dim, generator=g) * 0.5).to(self.device)
        if feature_hw is None:
            feature_hw = (max(1, round(height * 90 / 480)), max(1, round(width * 120 / 640)))
        self.fh, self.fw = feature_hw
        self.images = torch.empty(n_frames, height * width, 3, device=self.device)
        self.depths = torch.empty(n_frames, height * width, device=self.device)
        self.semantics = torch.empty(n_frames, height * width, dtype=torch.long, device=self.device)
        self.features = torch.empty(n_frames, self.fh * self.fw, feature_dim, dtype=torch.float16, device=self.device)
        # lazy: frames are rendered the first time a batch draws from them (the CPU arm of bench.py touches 8 frames per
        # step; rendering all 300 on host cores up front would dominate its run time)
        self.label_frac = label_frac
        self._have = None if not lazy else set()
        if not lazy:
            for i in range(n_frames):
                self._render_frame(i, label_frac)

    def _ensure(self, frames):
        if self._have is None:
            return
        for i in set(int(f) for f in frames) - self._have:
            self._render_frame(i, self.label_frac)
            self._have.add(i)

    # ------------------------------------------------------------ construction
    def _make_poses(self, g):
        poses = []
        for i in range(self.n):
            ang = 2 * math.pi * i / self.n
            rad = 0.7 + 0.3 * torch.rand(1, generator=g).item()
            eye = torch.tensor([rad * math.cos(ang), -0.2 + 0.5 * torch.rand(1, generator=g).item(), rad * math.sin(ang)])
            target = torch.tensor([0.12, -0.9, -0.18]) + 0.36 * (torch.rand(3, generator=g) - 0.5)
            fwd = target - eye
            fwd = fwd / fwd.norm()
            up = torch.tensor([0.0, 1.0, 0.0])
            right = torch.linalg.cross(fwd, up)
            right = right / right.norm()
            down = torch.linalg.cross(fwd, right)
            c2w = torch.eye(4)
            c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, down, fwd, eye
            poses.append(c2w)
        return torch.stack(poses)

    def camera_dirs(self, xs, ys):
        """Pixel coordinates (float) -> unnormalised camera-frame directions and their norms
        (autolabel/dataset.py:17-37)."""
        dirs = torch.stack([(xs - self.cx) / self.fx, (ys - self.cy) / self.fy, torch.ones_like(xs)], dim=1)
        norm = dirs.norm(dim=1, keepdim=True)
        return dirs / norm, norm

    def _rays(self, frame, xs, ys):
        d_cam, norm = self.camera_dirs(xs, ys)
        R = self.poses[frame, :3, :3]
        d = d_cam @ R.t()
        o = self.poses[frame, :3, 3].expand_as(d)
        return o.contiguous(), d.contiguous(), norm

    def _features_at(self, p, normal, obj, noise=True):
        cls = (obj > 0).float()
        v = torch.cat([p / ROOM, normal, cls[:, None], torch.ones_like(cls)[:, None]], dim=1)
        f = v @ self.proj
        if noise:
            f = f + 0.05 * torch.randn(f.shape, generator=self.gen, device=self.device)
        return torch.relu(f)          # autoencoded DINO codes are ReLU outputs (models.py:272-281)

    def _render_frame(self, i, label_frac):
        ys, xs = torch.meshgrid(torch.arange(self.h, device=self.device, dtype=torch.float32),
                                torch.arange(self.w, device=self.device, dtype=torch.float32), indexing='ij')
        o, d, norm = self._rays(i, xs.reshape(-1) + 0.5, ys.reshape(-1) + 0.5)
        t, normal, obj = _intersect(o, d)
        p = o + t[:, None] * d
        self.images[i] = _shade(p, normal, obj)
        self.depths[i] = t / norm[:, 0]
        sem = (obj > 0).long() if self.C == 2 else (obj % self.C)
        keep = torch.rand(sem.shape, generator=self.gen, device=self.device) < label_frac
        self.semantics[i] = torch.where(keep, sem, torch.full_like(sem, -1))
        fy, fx = torch.meshgrid(torch.arange(self.fh, device=self.device, dtype=torch.float32),
                                torch.arange(self.fw, device=self.device, dtype=torch.float32), indexing='ij')
        px = (fx.reshape(-1) + 0.5) * self.w / self.fw
        py = (fy.reshape(-1) + 0.5) * self.h / self.fh
        o2, d2, _ = self._rays(i, px, py)
        t2, n2, obj2 = _intersect(o2, d2)
        self.features[i] = self._features_at(o2 + t2[:, None] * d2, n2, obj2).half()

    # ------------------------------------------------------------ sampling
    @torch.no_grad()
    def next_train(self, batch_size=4096):
        """One training batch on the device, same dict as autolabel/dataset.py:182-242."""
        chunks = batch_size // self.chunk
        n = chunks * self.chunk
        dev = self.device
        frames = torch.randint(0, self.n, (chunks,), generator=self.gen, device=dev)
        self._ensure(frames.tolist() if self._have is not None else ())
        frames = frames.repeat_interleave(self.chunk)
        pix = torch.randint(0, self.h * self.w, (n,), generator=self.gen, device=dev)
        xs = (pix % self.w).float() + torch.rand(n, generator=self.gen, device=dev)
        ys = (pix // self.w).float() + torch.rand(n, generator=self.gen, device=dev)
        d_cam, norm = self.camera_dirs(xs, ys)
        R = self.poses[frames, :3, :3]
        rays_d = torch.bmm(R, d_cam[:, :, None])[:, :, 0].contiguous()
        rays_o = self.poses[frames, :3, 3].contiguous()
        fxi = ((pix % self.w).float() * self.fw / self.w).long().clamp(max=self.fw - 1)
        fyi = ((pix // self.w).float() * self.fh / self.h).long().clamp(max=self.fh - 1)
        return {
            'rays_o': rays_o, 'rays_d': rays_d, 'direction_norms': norm,
            'pixels': self.images[frames, pix], 'depth': self.depths[frames, pix],
            'semantic': self.semantics[frames, pix], 'features': self.features[frames, fyi * self.fw + fxi].float(),
        }

    @torch.no_grad()
    def get_test(self, i):
        """Full-frame rays of image i, same dict as autolabel/dataset.py:244-266 (tensors on the device)."""
        self._ensure([i])
        ys, xs = torch.meshgrid(torch.arange(self.h, device=self.device, dtype=torch.float32),
                                torch.arange(self.w, device=self.device, dtype=torch.float32), indexing='ij')
        o, d, norm = self._rays(i, xs.reshape(-1) + 0.5, ys.reshape(-1) + 0.5)
        return {
            'pixels': self.images[i].view(self.h, self.w, 3), 'rays_o': o.view(self.h, self.w, 3),
            'rays_d': d.view(self.h, self.w, 3), 'depth': self.depths[i].view(self.h, self.w),
            'semantic': self.semantics[i].view(self.h, self.w), 'H': self.h, 'W': self.w,
            'direction_norms': norm, 'features': self.features[i],
        }

    def bound(self):
        """create_model's bound rule (autolabel/model_utils.py:61-63)."""
        extents = self.max_bounds - self.min_bounds
        return float((extents - (self.min_bounds + self.max_bounds) * 0.5).max())
