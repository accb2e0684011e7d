"""``tinycudann``-shaped modules (the subset autolabel uses) on the sm_100a kernels.

The reference imports ``tinycudann as tcnn`` (autolabel/models.py:10) for
``tcnn.Encoding(n_input_dims, encoding_config)`` (Frequency :19-22,34-38, Grid/Hash :39-48,
SphericalHarmonics :97-101) and ``tcnn.Network(n_input_dims, n_output_dims, network_config)``
(FullyFusedMLP :84-92,104-113,127-136 and CutlassMLP :117-126).  ``sys.modules['tinycudann'] =
autolabel_b200.tcnn`` lets the unmodified reference model run on these kernels.

Conventions (tiny-cuda-nn is not in the reference tree and not pinned — see DESIGN.md "oracle"):
inputs are padded with ones to a multiple of 16, outputs to a multiple of 16, no biases; all
arithmetic enters the tensor cores as fp16 and accumulates in fp32; modules return fp32.
"""
import math

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib
from ._lib import call, ptr, stream_ptr
from .gridencoder import grid_encode, level_offsets


def _pad16(n):
    return (n + 15) // 16 * 16


# ------------------------------------------------------------------ Network
class _MlpFn(Function):

    @staticmethod
    def forward(ctx, x, params, dims):
        in_pad, hidden, out_pad, n_hidden, n_in, n_out = dims
        if not x.is_cuda:
            raise RuntimeError("tcnn.Network needs CUDA tensors; there is no CPU fallback")
        n = x.shape[0]
        dev = x.device
        xh = torch.ones(n, in_pad, dtype=torch.float16, device=dev)
        xh[:, :n_in] = x
        y = torch.empty(n, n_out, dtype=torch.float32, device=dev)
        p = params.float().contiguous()
        ctx.wide = _lib.lib.al_mlp_num_params(in_pad, hidden, out_pad, n_hidden) < 0
        ws = None
        if ctx.wide:
            # widths beyond the weight-resident fused kernels: tiled tcgen05 GEMMs, fp16 activations in a workspace
            training = 1 if (params.requires_grad or x.requires_grad) else 0
            ws = torch.empty(_lib.lib.al_mlp_wide_workspace(in_pad, hidden, out_pad, n_hidden, n, training),
                             dtype=torch.uint8, device=dev)
            call("al_mlp_wide_forward", in_pad, hidden, out_pad, n_hidden, ptr(p), ptr(xh), in_pad, n, None,
                 ptr(y), n_out, 0, 0, n_out, 0,
                 None, 0, 0, 0, 0, 0,
                 None, 0, 0, 0, 0, 0, ptr(ws), stream_ptr(dev))
        else:
            call("al_mlp_forward", in_pad, hidden, out_pad, n_hidden, ptr(p), ptr(xh), in_pad, n, None,
                 ptr(y), n_out, 0, 0, n_out, 0,
                 None, 0, 0, 0, 0, 0,
                 None, 0, 0, 0, 0, 0, stream_ptr(dev))
        ctx.save_for_backward(xh, p)
        ctx.ws = ws
        ctx.dims = dims
        ctx.needs_dx = x.requires_grad
        return y

    @staticmethod
    def backward(ctx, gy):
        xh, p = ctx.saved_tensors
        in_pad, hidden, out_pad, n_hidden, n_in, n_out = ctx.dims
        n = xh.shape[0]
        dev = xh.device
        gy = gy.float().contiguous()
        amax = gy.abs().amax().reshape(1).float()
        gp = torch.zeros_like(p)
        gx = torch.empty(n, n_in, dtype=torch.float32, device=dev) if ctx.needs_dx else None
        if ctx.wide:
            call("al_mlp_wide_backward", in_pad, hidden, out_pad, n_hidden, ptr(p), ptr(xh), in_pad, n, None, ptr(gy),
                 n_out, 0, n_out, ptr(amax), ptr(gp), ptr(gx), n_in, 0, n_in, ptr(ctx.ws), stream_ptr(dev))
            ctx.ws = None
        else:
            call("al_mlp_backward", in_pad, hidden, out_pad, n_hidden, ptr(p), ptr(xh), in_pad, n, None, ptr(gy),
                 n_out, 0, n_out, ptr(amax), ptr(gp), ptr(gx), 0, n_in, 0, n_in, stream_ptr(dev))
        return gx, gp, None


class Network(nn.Module):
    """tcnn.Network: bias-free MLP, ReLU hidden activations, linear output."""

    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
        super().__init__()
        cfg = dict(network_config)
        if cfg.get("activation", "ReLU") != "ReLU":
            raise NotImplementedError("only ReLU hidden activations are implemented")
        self.output_activation = cfg.get("output_activation", "None")
        if self.output_activation not in ("None", "ReLU"):
            raise NotImplementedError(f"output_activation {self.output_activation}")
        self.n_input_dims = int(n_input_dims)
        self.n_output_dims = int(n_output_dims)
        self.hidden = int(cfg["n_neurons"])
        self.n_hidden = int(cfg["n_hidden_layers"])
        self.in_pad = _pad16(self.n_input_dims)
        self.out_pad = _pad16(self.n_output_dims)
        n = _lib.lib.al_mlp_num_params(self.in_pad, self.hidden, self.out_pad, self.n_hidden)
        if n < 0:     # not a weight-resident fused shape: the tiled GEMM path (csrc/gemm_tc.cu)
            n = _lib.lib.al_mlp_wide_num_params(self.in_pad, self.hidden, self.out_pad, self.n_hidden)
        if n < 0:
            raise NotImplementedError(
                f"MLP shape in={self.in_pad} hidden={self.hidden} out={self.out_pad} n_hidden={self.n_hidden}: "
                "neither a fused shape (csrc/mlp_tc.cu AL_TC_CONFIGS) nor a wide shape (hidden multiple of 64, "
                "1 or 2 hidden layers, csrc/gemm_tc.cu)")
        self.params = nn.Parameter(torch.empty(n))
        self.seed = seed
        self.reset_parameters()

    @property
    def dims(self):
        return (self.in_pad, self.hidden, self.out_pad, self.n_hidden, self.n_input_dims, self.n_output_dims)

    def layer_shapes(self):
        shapes = [(self.hidden, self.in_pad)]
        if self.n_hidden == 2:
            shapes.append((self.hidden, self.hidden))
        shapes.append((self.out_pad, self.hidden))
        return shapes

    def reset_parameters(self):
        """Xavier-uniform per layer (tcnn's default initialisation)."""
        g = torch.Generator().manual_seed(self.seed)
        chunks = []
        for (o, i) in self.layer_shapes():
            s = math.sqrt(6.0 / (o + i))
            chunks.append((torch.rand(o * i, generator=g) * 2 - 1) * s)
        with torch.no_grad():
            self.params.copy_(torch.cat(chunks))

    def forward(self, x):
        prefix = x.shape[:-1]
        y = _MlpFn.apply(x.reshape(-1, self.n_input_dims).float(), self.params, self.dims)
        if self.output_activation == "ReLU":
            y = torch.relu(y)
        return y.view(*prefix, self.n_output_dims)


# ------------------------------------------------------------------ Encoding
class Encoding(nn.Module):
    """tcnn.Encoding for otype in {Frequency, SphericalHarmonics, Grid (Hash, Linear)}."""

    def __init__(self, n_input_dims, encoding_config, seed=1337):
        super().__init__()
        cfg = dict(encoding_config)
        self.otype = cfg["otype"]
        self.n_input_dims = int(n_input_dims)
        if self.otype == "Frequency":
            self.n_frequencies = int(cfg.get("n_frequencies", 12))
            self.n_output_dims = self.n_input_dims * 2 * self.n_frequencies
        elif self.otype == "SphericalHarmonics":
            if int(cfg.get("degree", 4)) != 4 or self.n_input_dims != 3:
                raise NotImplementedError("SphericalHarmonics: only degree 4 on 3-D inputs")
            self.n_output_dims = 16
        elif self.otype in ("Grid", "HashGrid"):
            if cfg.get("type", "Hash") != "Hash" or cfg.get("interpolation", "Linear") != "Linear":
                raise NotImplementedError("Grid: only type=Hash, interpolation=Linear")
            self.n_levels = int(cfg.get("n_levels", 16))
            self.n_features_per_level = int(cfg.get("n_features_per_level", 2))
            self.log2_hashmap_size = int(cfg.get("log2_hashmap_size", 19))
            self.base_resolution = int(cfg.get("base_resolution", 16))
            self.per_level_scale = float(cfg.get("per_level_scale", 2.0))
            offsets = level_offsets(self.n_input_dims, self.n_levels, self.per_level_scale, self.base_resolution,
                                    self.log2_hashmap_size)
            self.register_buffer("offsets", torch.from_numpy(offsets))
            self.params = nn.Parameter(torch.empty(int(offsets[-1]) * self.n_features_per_level))
            g = torch.Generator().manual_seed(seed)
            with torch.no_grad():
                self.params.copy_((torch.rand(self.params.numel(), generator=g) * 2 - 1) * 1e-4)
            self.n_output_dims = self.n_levels * self.n_features_per_level
        else:
            raise NotImplementedError(f"encoding otype {self.otype}")

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError("tcnn.Encoding needs CUDA tensors; there is no CPU fallback")
        prefix = x.shape[:-1]
        x = x.reshape(-1, self.n_input_dims).float().contiguous()
        B = x.shape[0]
        dev = x.device
        if self.otype == "Frequency":
            out = torch.empty(B, self.n_output_dims, dtype=torch.float32, device=dev)
            call("al_freq_encode", ptr(x), B, self.n_input_dims, self.n_frequencies, ptr(out), stream_ptr(dev))
        elif self.otype == "SphericalHarmonics":
            out = torch.empty(B, 16, dtype=torch.float32, device=dev)
            call("al_sh_encode", ptr(x), B, ptr(out), stream_ptr(dev))
        else:
            out = grid_encode(x, self.params.view(-1, self.n_features_per_level), self.offsets,
                              self.per_level_scale, self.base_resolution, False, 0)
        return out.view(*prefix, self.n_output_dims)
