"""Pins oracle/postprocess_oracle.py (CPU, numpy) on the libraries the reference itself calls for these steps:
sklearn's PCA.transform (scripts/render.py:63) and torch.argmax / torch.norm (scripts/render.py:72-80)."""
import numpy as np
import torch

from oracle import postprocess_oracle as po


def _data(n=500, F=64, T=7, seed=0):
    rng = np.random.RandomState(seed)
    return rng.normal(size=(n, F)).astype(np.float32), rng.normal(size=(T, F)).astype(np.float32)


def test_pca_matches_sklearn_transform():
    from sklearn.decomposition import PCA
    feats, _ = _data(2000, 64)
    pca = PCA(n_components=3).fit(feats)
    ref = pca.transform(feats)
    got = po.pca_project(feats, pca.mean_, pca.components_)
    assert np.abs(ref - got).max() < 1e-4
    fmin, frange = ref.min(0), ref.max(0) - ref.min(0)
    col = po.pca_colors(feats, pca.mean_, pca.components_, fmin, frange)
    want = (np.clip((ref - fmin) / frange, 0., 1.) * 255.).astype(np.uint8)
    assert (np.abs(col.astype(int) - want.astype(int)) <= 1).all() and col.dtype == np.uint8


def test_text_argmax_matches_the_reference_loop():
    feats, text = _data(480, 64, 9, seed=1)
    H, W = 24, 20
    f = torch.from_numpy(feats).view(H, W, 64)
    t = torch.from_numpy(text)
    f = f / torch.norm(f, dim=-1, keepdim=True)                     # scripts/render.py:72-80, verbatim shape logic
    sims = torch.zeros((H, W, t.shape[0]))
    for i in range(H):
        sims[i, :, :] = (f[i, :, None] * t).sum(dim=-1)
    assert np.array_equal(sims.argmax(dim=-1).numpy().reshape(-1), po.text_argmax(feats, text))


def test_semantic_argmax_first_maximum():
    logits = np.array([[0.1, 0.7, 0.7], [2.0, 2.0, 1.0], [-1.0, -3.0, -1.0]], np.float32)
    assert po.semantic_argmax(logits).tolist() == torch.from_numpy(logits).argmax(-1).tolist() == [1, 0, 0]
    assert po.rgb_u8(np.array([0.0, 0.5, 1.0], np.float32)).tolist() == [0, 127, 255]
