"""Seeded synthetic inputs shared by the parity tests (numpy, device independent)."""
import numpy as np


def make_rays(n, bound, seed=0, inside=True):
    """Origins inside (or around) the cube, unit directions; a few rays miss the box on purpose."""
    rng = np.random.RandomState(seed)
    o = rng.uniform(-0.8 * bound, 0.8 * bound, size=(n, 3)).astype(np.float32)
    if not inside:
        o[: n // 4] = rng.uniform(1.2 * bound, 2.0 * bound, size=(n // 4, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d.astype(np.float32)


def make_density_grid(cascade, H=128, seed=0, fill=0.15):
    """Blobby occupancy: a few Gaussian blobs per cascade in Morton order is overkill for a test;
    use random smooth-ish values with a fraction `fill` above the 0.01 threshold, plus -1 cells."""
    rng = np.random.RandomState(seed)
    g = rng.uniform(0, 1, size=(cascade, H ** 3)).astype(np.float32)
    # make occupancy spatially coherent in Morton order: blocks of 512 consecutive cells (8x8x8 bricks)
    brick = rng.uniform(0, 1, size=(cascade, H ** 3 // 512)).astype(np.float32)
    occ = np.repeat(brick < fill, 512, axis=1)
    g = np.where(occ, 0.02 + g, 0.001 * g).astype(np.float32)
    g[:, ::97] = -1.0
    return g


def aabb_of(bound):
    return np.array([-bound, -bound, -bound, bound, bound, bound], dtype=np.float32)


# ------------------------------------------------------------------ parity report
import json
import os

_REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_report.json")


def record(name, **values):
    """Append measured parity errors to gpurun_out/parity_report.json (copied to profiles/ per round)."""
    try:
        os.makedirs(os.path.dirname(_REPORT), exist_ok=True)
        data = json.load(open(_REPORT)) if os.path.exists(_REPORT) else {}
        data[name] = {k: (float(v) if not isinstance(v, (str, int)) else v) for k, v in values.items()}
        json.dump(data, open(_REPORT, "w"), indent=1, sort_keys=True)
    except Exception:
        pass


def rel_max(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def rel_l2(a, b):
    return ((a - b).double().norm() / (b.double().norm() + 1e-30)).item()


# ------------------------------------------------------------------ replayable RNG for update_extra_state parity
class TorchRngTape:
    """Stands in for the `torch` module inside a renderer module: everything is delegated to torch except the two random
    draws NeRFRenderer.update_extra_state makes (torch.rand_like, torch.randint), which are served from a seeded numpy
    stream.  CPU and CUDA generators of torch produce different streams for the same seed; with this tape the
    reference's code (on CPU) and the product's code (on the GPU) see the SAME numbers as long as they ask for them in
    the same order with the same shapes — which is exactly the control-flow property the golden test pins."""

    def __init__(self, torch_module, seed):
        self._t = torch_module
        self._rs = np.random.RandomState(seed)
        self.calls = []

    def __getattr__(self, name):
        return getattr(self._t, name)

    def rand_like(self, t, **kw):
        self.calls.append(("rand_like", tuple(t.shape)))
        v = self._rs.random_sample(tuple(t.shape)).astype(np.float32)
        return self._t.from_numpy(v).to(t.device)

    def randint(self, low, high, size, dtype=None, device=None, **kw):
        size = tuple(int(s) for s in size)
        self.calls.append(("randint", int(low), int(high), size))
        v = self._rs.randint(int(low), int(high), size=size).astype(np.int64)
        out = self._t.from_numpy(v)
        if device is not None:
            out = out.to(device)
        return out if dtype is None else out.to(dtype)


def checker_density(xyz):
    """An exactly representable density field (integers times 4 in fp32, identical on CPU and GPU):
    4 * ((floor(4x) + floor(4y) + floor(4z)) mod 5)."""
    import torch
    s = torch.floor(xyz * 4.0).sum(dim=-1)
    return torch.remainder(s, 5.0) * 4.0
