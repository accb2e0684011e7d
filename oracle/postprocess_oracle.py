"""CPU restatement (numpy) of the reference's per-frame render epilogues.  TEST INFRASTRUCTURE ONLY.

  semantic_argmax   scripts/export.py:78-90          outputs['semantic'].argmax(dim=-1)
  text_argmax       scripts/render.py:69-82,         features / norm, dot with the text features row by row, argmax
                    autolabel/evaluation.py:295-318
  pca_colors        scripts/render.py:61-66          sklearn PCA.transform (no whitening): (X - mean_) @ components_.T,
                                                     clip((x - min) / range, 0, 1) * 255 as uint8
  rgb_u8            scripts/render.py:104            (image * 255).astype(uint8)

Pinned by tests/test_oracle_postprocess_cpu.py against sklearn's own PCA.transform and torch.argmax (the reference
holds no golden vectors for these steps).
"""
import numpy as np


def semantic_argmax(logits):
    return np.argmax(logits, axis=-1).astype(np.int32)


def text_similarities(features, text_features):
    f = features / np.linalg.norm(features, axis=-1, keepdims=True)
    return (f[..., None, :] * text_features).sum(axis=-1)


def text_argmax(features, text_features):
    return np.argmax(text_similarities(features, text_features), axis=-1).astype(np.int32)


def pca_project(features, mean, components):
    return (features - mean) @ components.T


def pca_colors(features, mean, components, fmin, frange):
    x = pca_project(features.astype(np.float32), mean.astype(np.float32), components.astype(np.float32))
    x = np.clip((x - fmin) / frange, 0.0, 1.0)
    return (x * 255.0).astype(np.uint8)


def rgb_u8(image):
    return (image * 255.0).astype(np.uint8)
