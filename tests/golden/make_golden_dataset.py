"""Freeze outputs of the reference's OWN batch sampler (autolabel/dataset.py: `_compute_direction`
:17-37, `BaseDataset._next_train` :182-242, `_get_test` :244-266) into tests/golden/ref_dataset.npz.

Runs in the dev container (CPU; needs /root/reference + numba):  python tests/golden/make_golden_dataset.py
The reference module is imported unmodified; absent third-party modules that the sampler never calls
(h5py, torch_ngp's optional deps) are stubbed.  A `BaseDataset` is filled with small seeded arrays of the
kinds `SceneDataset._load_images` / `_load_features` produce (images fp32 [n,HW,3], depths uint16 mm,
semantics uint8, features fp16 [n, fh*fw, F], poses fp32).  The random draws of `_next_train` are
recorded by wrapping `_compute_direction` (image index + ray indices per chunk); the sub-pixel jitter
lives inside the numba function, so the golden batch is produced with `randomize=False` (pixel centres)
and a second, jittered call only records the jitter it implies for range checks.
"""
import os
import random
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("AUTOLABEL_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "ref_dataset.npz")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__file__ = f"<stub {name}>"
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference_dataset():
    import torch  # noqa: F401  (first: torch.library inspects sys.modules)
    sys.path.insert(0, REF)
    if "h5py" not in sys.modules:
        try:
            import h5py  # noqa: F401
        except ImportError:
            _stub("h5py")
    # autolabel.utils pulls Scene helpers only; torch_ngp.nerf.provider pulls the NeRF dataset with optional deps
    for name in ("trimesh", "mcubes", "tensorboardX", "torch_ema", "torch_scatter"):
        if name not in sys.modules:
            try:
                __import__(name)
            except ImportError:
                _stub(name)
    try:
        import torch_ngp.nerf.provider  # noqa: F401
    except Exception:
        # the sampler needs nerf_matrix_to_ngp only at scene-loading time (never called here)
        _stub("torch_ngp"); _stub("torch_ngp.nerf")
        _stub("torch_ngp.nerf.provider", nerf_matrix_to_ngp=lambda pose, scale=1.0: pose)
    from autolabel import dataset
    return dataset


class _Camera:
    def __init__(self, w, h, fx, fy, cx, cy):
        self.size = (w, h)
        self.camera_matrix = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
        self.fx, self.fy, self.cx, self.cy = fx, fy, cx, cy


def make_arrays(seed=0, n=5, w=40, h=30, fw=10, fh=8, F=16):
    rng = np.random.RandomState(seed)
    images = rng.uniform(0, 1, size=(n, h * w, 3)).astype(np.float32)
    depths = rng.randint(0, 6000, size=(n, h * w)).astype(np.uint16)
    semantics = (rng.randint(0, 3, size=(n, h * w)) * (rng.uniform(size=(n, h * w)) < 0.2)).astype(np.uint8)
    features = rng.normal(size=(n, fh * fw, F)).astype(np.float16)
    poses = np.zeros((n, 4, 4), dtype=np.float32)
    for i in range(n):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        poses[i, :3, :3] = q.astype(np.float32)
        poses[i, :3, 3] = rng.uniform(-1, 1, size=3).astype(np.float32)
        poses[i, 3, 3] = 1
    return dict(images=images, depths=depths, semantics=semantics, features=features, poses=poses,
                w=w, h=h, fw=fw, fh=fh, F=F, fx=0.8 * w + 0.37, fy=0.8 * w - 0.21, cx=w / 2.0 + 0.3, cy=h / 2.0 - 0.4)


def build_reference_dataset(dataset, a, batch_size):
    cam = _Camera(a['w'], a['h'], a['fx'], a['fy'], a['cx'], a['cy'])
    ds = dataset.BaseDataset(batch_size, cam)
    ds.index_sampler = dataset.IndexSampler()
    ds.images, ds.depths, ds.semantics = a['images'], a['depths'], a['semantics']
    ds.index_sampler.update(ds.semantics)
    ds.poses = a['poses']
    ds.rotations = np.ascontiguousarray(a['poses'][:, :3, :3])
    ds.origins = a['poses'][:, :3, 3]
    ds.n_examples = a['images'].shape[0]
    ds.pixel_indices = np.arange(a['w'] * a['h'])[::1]
    ds.features = a['features']
    ds.feature_width, ds.feature_height, ds.feature_dim = a['fw'], a['fh'], a['F']
    scale_factor = np.array([a['fw'] / cam.size[0], a['fh'] / cam.size[1]])
    ds._scale_to_feature_xy = lambda xy: (xy * scale_factor).astype(int)
    return ds


def main():
    dataset = import_reference_dataset()
    a = make_arrays()
    batch = 4 * 512
    ds = build_reference_dataset(dataset, a, batch)
    rec = {'image_index': [], 'ray_indices': []}
    orig = ds._compute_direction

    def centred(image_index, ray_indices, randomize=False):
        rec['image_index'].append(int(image_index))
        rec['ray_indices'].append(np.asarray(ray_indices).astype(np.int64))
        return orig(image_index, ray_indices, randomize=False)

    random.seed(3)
    np.random.seed(4)
    ds._compute_direction = centred
    out = ds._next_train()
    ds._compute_direction = orig
    test = ds._get_test(2)
    # a jittered call of the numba function itself (its own RNG): only ranges are checked against it
    d_j, n_j = orig(1, rec['ray_indices'][0], randomize=True)

    np.savez_compressed(
        OUT, images=a['images'], depths=a['depths'], semantics=a['semantics'], features=a['features'], poses=a['poses'],
        meta=np.array([a['w'], a['h'], a['fw'], a['fh'], a['F']], dtype=np.int64),
        intrinsics=np.array([a['fx'], a['fy'], a['cx'], a['cy']], dtype=np.float64),
        image_index=np.array(rec['image_index'], dtype=np.int32), ray_indices=np.concatenate(rec['ray_indices']).astype(np.int32),
        train_rays_o=out['rays_o'], train_rays_d=out['rays_d'], train_norms=out['direction_norms'],
        train_pixels=out['pixels'], train_depth=out['depth'], train_semantic=out['semantic'].astype(np.int64),
        train_features=out['features'],
        test_rays_o=test['rays_o'], test_rays_d=test['rays_d'], test_norms=test['direction_norms'],
        test_depth=test['depth'].astype(np.float64), test_semantic=test['semantic'].astype(np.int64),
        jit_rays_d=d_j, jit_norms=n_j)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", "chunks", rec['image_index'])


if __name__ == "__main__":
    main()
