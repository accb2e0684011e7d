"""ALNetwork — same constructor, attributes and methods as the reference's
``autolabel/models.py`` (:62-265), on the sm_100a kernels.

``forward / density / color / semantic`` run through the per-module operators (autograd works,
any batch shape); ``render()`` with ``cuda_ray=True`` uses the fused field pipeline
(``renderer._FusedRender``).  Parameters: ``encoder`` (hash table, if any), ``sigma_net``,
``color_net``, ``semantic_features``, ``semantic_out`` — flat fp32 ``params`` vectors like tcnn's.
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, tcnn
from ._lib import FieldDesc, call, ptr, stream_ptr
from .gridencoder import GridEncoder
from .renderer import NeRFRenderer


class _TruncExp(torch.autograd.Function):
    """torch_ngp/activation.py:1-17: exp forward, exp(clamp(x, -15, 15)) backward."""

    @staticmethod
    def forward(ctx, x):
        x = x.float()
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _TruncExp.apply


class FreqEncoder(nn.Module):
    """models.py:15-29: Frequency(10) of the position normalised to [0,1]."""

    def __init__(self, input_dim):
        super().__init__()
        self.encoder = tcnn.Encoding(input_dim, {"otype": "Frequency", "n_frequencies": 10})
        self.n_output_dims = self.encoder.n_output_dims

    def forward(self, x, bound):
        return self.encoder((x + bound) / (2.0 * bound))


class HGFreqEncoder(nn.Module):
    """models.py:31-59: Frequency(2) of the raw position ++ hash grid of the clipped normalised one."""

    def __init__(self, input_dim):
        super().__init__()
        self.encoder = tcnn.Encoding(input_dim, {"otype": "Frequency", "n_frequencies": 2})
        self.grid_encoding = tcnn.Encoding(input_dim, {
            "otype": "Grid", "type": "Hash", "n_levels": 16, "n_features_per_level": 2,
            "log2_hashmap_size": 19, "base_resolution": 16, "per_level_scale": 2.0, "interpolation": "Linear"})
        self.n_output_dims = self.encoder.n_output_dims + self.grid_encoding.n_output_dims

    def forward(self, x, bound):
        freq = self.encoder(x)
        normalized = torch.clip((x + bound) / (2.0 * bound), 0.0, 1.0)
        return torch.cat([freq, self.grid_encoding(normalized)], dim=-1)


class _HGEncoder(GridEncoder):
    """get_encoder('hashgrid', desired_resolution=2**18) of models.py:142-143."""

    def forward(self, x, bound=1):
        return super().forward(x, bound=bound)


class ALNetwork(NeRFRenderer):

    def __init__(self, encoding='hg', num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=2,
                 hidden_dim_color=64, hidden_dim_semantic=64, semantic_classes=2, bound=1, **kwargs):
        super().__init__(bound, **kwargs)
        if geo_feat_dim != 15:
            raise NotImplementedError("geo_feat_dim must be 15 (1 + 15 = one 16-wide MLP output tile)")
        # The reference signature defaults num_layers_color to 3; its only caller (create_model, model_utils.py:61-74)
        # passes 2, which is the default here so that ALNetwork() constructs (documented in INTEGRATION.md).
        if num_layers != 2 or num_layers_color != 2:
            raise NotImplementedError(
                "this build instantiates 2-hidden-layer density / colour MLPs (what autolabel's create_model "
                "uses, model_utils.py:61-74); pass num_layers=2, num_layers_color=2")
        self.encoding = encoding
        self.num_layers = num_layers
        self.hidden_dim = hidden_dim
        self.geo_feat_dim = geo_feat_dim
        self.encoder, self.in_dim = self._get_encoder(encoding)
        ffmlp = {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": "None"}
        self.sigma_net = tcnn.Network(self.in_dim, 1 + geo_feat_dim,
                                      dict(ffmlp, n_neurons=hidden_dim, n_hidden_layers=num_layers), seed=11)
        self.num_layers_color = num_layers_color
        self.hidden_dim_color = hidden_dim_color
        self.encoder_dir = tcnn.Encoding(3, {"otype": "SphericalHarmonics", "degree": 4})
        self.color_features = self.encoder_dir.n_output_dims + geo_feat_dim
        self.color_net = tcnn.Network(self.color_features, 3,
                                      dict(ffmlp, n_neurons=hidden_dim_color, n_hidden_layers=num_layers_color),
                                      seed=12)
        self.hidden_dim_semantic = hidden_dim_semantic
        self.semantic_classes = semantic_classes
        self.semantic_features = tcnn.Network(geo_feat_dim, hidden_dim_semantic,
                                              dict(ffmlp, otype="CutlassMLP", n_neurons=hidden_dim_semantic,
                                                   n_hidden_layers=2), seed=13)
        self.semantic_out = tcnn.Network(hidden_dim_semantic + geo_feat_dim, semantic_classes,
                                         dict(ffmlp, n_neurons=64, n_hidden_layers=1), seed=14)
        self._desc_keepalive = None

    def _get_encoder(self, encoding):
        if encoding == 'freq':
            enc = FreqEncoder(3)
            return enc, enc.n_output_dims
        if encoding == 'hg':
            enc = _HGEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                             desired_resolution=2 ** 18, gridtype='hash')
            return enc, enc.output_dim
        if encoding == 'hg+freq':
            enc = HGFreqEncoder(3)
            return enc, enc.n_output_dims
        raise NotImplementedError(f"Unknown input encoding {encoding}")

    # ------------------------------------------------------------ module-level API (models.py:150-256)
    def forward(self, x, d):
        x = self.encoder(x, bound=self.bound)
        h = self.sigma_net(x)
        sigma = trunc_exp(h[..., 0])
        geo_feat = F.relu(h[..., 1:])
        d = self.encoder_dir(d)
        rgb = torch.sigmoid(self.color_net(torch.cat([d, geo_feat], dim=-1)))
        features = self.semantic_features(geo_feat)
        semantic = self.semantic_out(torch.cat([F.relu(features), geo_feat], dim=-1))
        return sigma, rgb, F.softmax(semantic, dim=-1)

    def density(self, x):
        x = self.encoder(x, bound=self.bound)
        h = self.sigma_net(x)
        return {'sigma': trunc_exp(h[..., 0]), 'geo_feat': h[..., 1:]}

    def color(self, x, d, mask=None, geo_feat=None, **kwargs):
        if mask is not None:
            rgbs = torch.zeros(mask.shape[0], 3, dtype=x.dtype, device=x.device)
            if not mask.any():
                return rgbs
            x, d, geo_feat = x[mask], d[mask], geo_feat[mask]
        d = self.encoder_dir((d + 1) / 2)
        h = torch.sigmoid(self.color_net(torch.cat([d, geo_feat], dim=-1)))
        if mask is not None:
            rgbs[mask] = h.to(rgbs.dtype)
            return rgbs
        return h

    def semantic(self, geo_features, sigma):
        sem_features = self.semantic_features(geo_features)
        features = torch.cat([F.relu(sem_features), geo_features], dim=1)
        return self.semantic_out(features), sem_features

    def get_params(self, lr):
        return [
            {'params': self.encoder.parameters(), 'lr': lr},
            {'params': self.sigma_net.parameters(), 'lr': lr},
            {'params': self.encoder_dir.parameters(), 'lr': lr},
            {'params': self.color_net.parameters(), 'lr': lr},
            {'params': self.semantic_features.parameters(), 'lr': lr},
            {'params': self.semantic_out.parameters(), 'lr': lr},
        ]

    def network_parameters(self):
        return (list(self.sigma_net.parameters()) + list(self.color_net.parameters()) +
                list(self.semantic_features.parameters()) + list(self.semantic_out.parameters()))

    # ------------------------------------------------------------ fused field plumbing
    @property
    def n_channels(self):
        return 3 + self.semantic_classes + self.hidden_dim_semantic

    def _table(self):
        if self.encoding == 'hg':
            return self.encoder.embeddings
        if self.encoding == 'hg+freq':
            return self.encoder.grid_encoding.params
        return None

    def _grid_meta(self):
        if self.encoding == 'hg':
            e = self.encoder
            return e.offsets, e.num_levels, float(np.log2(e.per_level_scale)), e.base_resolution
        if self.encoding == 'hg+freq':
            e = self.encoder.grid_encoding
            return e.offsets, e.n_levels, float(np.log2(e.per_level_scale)), e.base_resolution
        return None, 0, 0.0, 0

    def field_params(self):
        """(table | None, w_sigma, w_color, w_semf, w_semo)"""
        return (self._table(), self.sigma_net.params, self.color_net.params, self.semantic_features.params,
                self.semantic_out.params)

    def field_desc(self):
        d = FieldDesc()
        d.encoding = {'freq': 0, 'hg': 1, 'hg+freq': 2}[self.encoding]
        d.in_pad = self.sigma_net.in_pad
        d.hidden = self.hidden_dim
        d.hidden_color = self.hidden_dim_color
        d.feat_dim = self.hidden_dim_semantic
        d.n_classes = self.semantic_classes
        d.bound = float(self.bound)
        offsets, L, S, H = self._grid_meta()
        d.L, d.H, d.gridtype, d.S = int(L), int(H), 0, float(S)
        table = self._table()
        d.offsets = ptr(offsets)
        d.table = ptr(table)
        d.w_sigma = ptr(self.sigma_net.params)
        d.w_color = ptr(self.color_net.params)
        d.w_semf = ptr(self.semantic_features.params)
        d.w_semo = ptr(self.semantic_out.params)
        return d

    @torch.no_grad()
    def density_only(self, xyz):
        """sigma [n] of positions [n,3] with the fused encoder + density MLP (no heads)."""
        xyz = xyz.float().contiguous()
        n = xyz.shape[0]
        dev = xyz.device
        desc = self.field_desc()
        vals = torch.empty(n, 1, dtype=torch.float32, device=dev)
        ws = torch.empty(_lib.lib.al_field_workspace(ctypes.byref(desc), n, 0), dtype=torch.uint8, device=dev)
        call("al_field_forward", ctypes.byref(desc), ptr(xyz), None, None, n, None, ptr(vals), 1, None, 1, ptr(ws),
             stream_ptr(dev))
        return vals.view(-1)

    @torch.no_grad()
    def field_values(self, xyz, dirs):
        """[n, 1+3+C+F] = (sigma, rgb, logits, features) of the fused field (inference)."""
        xyz = xyz.float().contiguous()
        dirs = dirs.float().contiguous()
        n = xyz.shape[0]
        dev = xyz.device
        desc = self.field_desc()
        ldv = 1 + self.n_channels
        vals = torch.empty(n, ldv, dtype=torch.float32, device=dev)
        ws = torch.empty(_lib.lib.al_field_workspace(ctypes.byref(desc), n, 0), dtype=torch.uint8, device=dev)
        call("al_field_forward", ctypes.byref(desc), ptr(xyz), ptr(dirs), None, n, None, ptr(vals), ldv, None, 0,
             ptr(ws), stream_ptr(dev))
        return vals
