"""The numpy restatement of the reference's batch sampler (oracle/dataset_oracle.py) against outputs of the
reference's OWN code (autolabel/dataset.py, numba `_compute_direction` + `BaseDataset._next_train/_get_test`),
frozen in tests/golden/ref_dataset.npz by tests/golden/make_golden_dataset.py.  Runs without a GPU."""
import os

import numpy as np

from oracle import dataset_oracle as do

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_dataset.npz"))


def _scene():
    w, h, fw, fh, F = (int(v) for v in G['meta'])
    intr = tuple(float(v) for v in G['intrinsics'])
    R = np.ascontiguousarray(G['poses'][:, :3, :3])
    O = G['poses'][:, :3, 3]
    return w, h, fw, fh, F, intr, R, O


def test_next_train_matches_reference():
    w, h, fw, fh, F, intr, R, O = _scene()
    out = do.next_train(G['images'], G['depths'], G['semantics'], G['features'], R, O, w, fw, fh, h, intr,
                        G['image_index'], G['ray_indices'])
    for k, gk in [('rays_o', 'train_rays_o'), ('direction_norms', 'train_norms'), ('pixels', 'train_pixels'),
                  ('depth', 'train_depth'), ('semantic', 'train_semantic'), ('features', 'train_features')]:
        assert out[k].dtype == G[gk].dtype and np.array_equal(out[k], G[gk]), k      # bit-exact
    assert np.abs(out['rays_d'] - G['train_rays_d']).max() <= 1e-6                     # BLAS mat-vec order: 1 ulp
    assert out['semantic'].min() == -1 and (out['semantic'] >= 0).any()


def test_get_test_matches_reference():
    w, h, fw, fh, F, intr, R, O = _scene()
    o, d, nrm = do.get_test(R, O, w, h, intr, 2)
    assert np.array_equal(o, G['test_rays_o'])
    assert np.array_equal(nrm, G['test_norms'])
    assert np.abs(d - G['test_rays_d']).max() <= 1e-6


def test_jittered_directions_stay_inside_their_pixel():
    """The reference's jittered call (numba RNG, not reproducible here): every direction must be reachable by SOME
    jitter in [0,1)^2 of the same pixel, i.e. lie between the oracle's corner directions."""
    w, h, fw, fh, F, intr, R, O = _scene()
    idx = G['ray_indices'][:512]
    lo = np.zeros((512, 2), np.float32)
    hi = np.full((512, 2), np.float32(1.0 - 1e-7))
    R1 = R[1]
    cam = G['jit_rays_d'] @ R1                      # back to the camera frame (R orthonormal): d_cam = R^T d
    cam = cam * G['jit_norms']
    fx, fy, cx, cy = intr
    xs = cam[:, 0] * fx + cx
    ys = cam[:, 1] * fy + cy
    px, py = idx % w, idx // w
    assert np.all(xs >= px - 1e-3) and np.all(xs <= px + 1 + 1e-3)
    assert np.all(ys >= py - 1e-3) and np.all(ys <= py + 1 + 1e-3)
    d0, _ = do.compute_direction(R1, idx, w, fx, fy, cx, cy, lo)
    d1, _ = do.compute_direction(R1, idx, w, fx, fy, cx, cy, hi)
    assert np.isfinite(d0).all() and np.isfinite(d1).all()
