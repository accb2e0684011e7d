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
