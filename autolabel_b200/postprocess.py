"""Render epilogues of the export / render scripts on the device.

Mirrors what the reference computes per frame after ``model.render(...)``:

* ``scripts/export.py:78-90``  ``outputs['semantic'].argmax(dim=-1)``
* ``scripts/render.py:69-82`` / ``autolabel/evaluation.py:295-318``  ``compute_semantics``: features normalised, dot
  product with the encoded class prompts (a Python loop over the H rows in the reference), argmax
* ``scripts/render.py:40-66``  ``FeatureTransformer.__call__``: PCA projection of the feature map to 3 channels,
  ``clip((x - min) / range, 0, 1) * 255`` as uint8
* ``scripts/render.py:104``  ``(image * 255).astype(uint8)``

in ONE kernel launch (``al_render_epilogue``), so a frame goes back to the host as 14 bytes per pixel instead of the
full fp32 maps.  No CPU fallback: CUDA tensors only.
"""
import numpy as np
import torch

from ._lib import call, ptr, require_cuda, stream_ptr


class FeatureTransformer:
    """``scripts/render.py:40-66`` without the h5py / extractor plumbing: built from the arrays the reference reads
    from ``features.hdf`` (``attrs['pca']`` -> ``mean_`` / ``components_``, ``attrs['min']``, ``attrs['range']``) and
    the already encoded text features [T, F]."""

    def __init__(self, pca_mean, pca_components, feature_min, feature_range, text_features=None, device="cuda"):
        def dev(a):
            return None if a is None else torch.as_tensor(np.asarray(a, dtype=np.float32)).to(device).contiguous()
        self.pca_mean, self.pca_components = dev(pca_mean), dev(pca_components)
        self.feature_min, self.feature_range = dev(feature_min), dev(feature_range)
        self.text_features = dev(text_features) if not torch.is_tensor(text_features) else text_features.float().to(device).contiguous()
        assert self.pca_components.shape[0] == 3 and self.pca_components.shape[1] == self.pca_mean.shape[0]

    @classmethod
    def from_sklearn(cls, pca, feature_min, feature_range, text_features=None, device="cuda"):
        return cls(pca.mean_, pca.components_[:3], feature_min, feature_range, text_features, device)

    def __call__(self, p_features):
        """[H, W, F] feature map (device tensor) -> [H, W, 3] uint8 (device tensor)."""
        H, W, F = p_features.shape
        return render_epilogue({'semantic_features': p_features.reshape(H * W, F)}, feature_transform=self,
                               want=('pca8',))['pca8'].view(H, W, 3)


def render_epilogue(outputs, text_features=None, feature_transform=None, want=('rgb8', 'label', 'text_label', 'pca8')):
    """outputs: the dict of ``model.render`` (any leading shape).  Returns the requested maps with the leading shape of
    the inputs: rgb8 uint8 [..., 3], label int32 [...], text_label int32 [...], pca8 uint8 [..., 3]."""
    ft = feature_transform
    if text_features is None and ft is not None:
        text_features = ft.text_features
    image = outputs.get('image') if 'rgb8' in want else None
    logits = outputs.get('semantic') if 'label' in want else None
    need_feat = ('text_label' in want and text_features is not None) or ('pca8' in want and ft is not None)
    feat = outputs.get('semantic_features') if need_feat else None
    first = next(t for t in (image, logits, feat) if t is not None)
    require_cuda(first)
    dev = first.device
    lead = first.shape[:-1]
    N = int(np.prod(lead)) if len(lead) else 1

    def rows(t):          # [N, width] view with a row stride (no copy for column slices of the compositing buffer)
        if t is None:
            return None, 0, 0
        t = t.reshape(N, t.shape[-1])
        if t.dtype != torch.float32 or t.stride(1) != 1:
            t = t.float().contiguous()
        return t, t.stride(0), t.shape[1]
    image = None if image is None else image.reshape(N, 3).float().contiguous()
    logits, ld_logits, C = rows(logits)
    feat, ld_feat, F = rows(feat)
    res = {}
    rgb8 = torch.empty(N, 3, dtype=torch.uint8, device=dev) if image is not None else None
    label = torch.empty(N, dtype=torch.int32, device=dev) if logits is not None else None
    text_label = pca8 = None
    T = 0
    if feat is not None and 'text_label' in want and text_features is not None:
        text_features = text_features.to(dev).float().contiguous()
        T = text_features.shape[0]
        assert text_features.shape[1] == F
        text_label = torch.empty(N, dtype=torch.int32, device=dev)
    if feat is not None and 'pca8' in want and ft is not None:
        assert ft.pca_mean.shape[0] == F
        pca8 = torch.empty(N, 3, dtype=torch.uint8, device=dev)
    call("al_render_epilogue", ptr(image), ptr(logits), int(ld_logits), ptr(feat), int(ld_feat), N, int(C), int(F),
         ptr(text_features) if text_label is not None else None, int(T),
         ptr(ft.pca_mean) if pca8 is not None else None, ptr(ft.pca_components) if pca8 is not None else None,
         ptr(ft.feature_min) if pca8 is not None else None, ptr(ft.feature_range) if pca8 is not None else None,
         ptr(rgb8), ptr(label), ptr(text_label), ptr(pca8), stream_ptr(dev))
    if rgb8 is not None:
        res['rgb8'] = rgb8.view(*lead, 3)
    if label is not None:
        res['label'] = label.view(*lead)
    if text_label is not None:
        res['text_label'] = text_label.view(*lead)
    if pca8 is not None:
        res['pca8'] = pca8.view(*lead, 3)
    return res


def compute_semantics(outputs, classes, feature_transform):
    """``scripts/render.py:69-82``: open-vocabulary labels when class prompts are given, else the argmax of the
    semantic head.  Returns a device int32 tensor with the leading shape of the maps."""
    if classes is not None:
        return render_epilogue(outputs, feature_transform=feature_transform, want=('text_label',))['text_label']
    return render_epilogue(outputs, want=('label',))['label']
