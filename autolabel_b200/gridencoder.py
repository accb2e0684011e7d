"""Multiresolution hash-grid encoder — same module interface as the reference's
``torch_ngp/gridencoder/grid.py`` (``GridEncoder`` :91-156, ``grid_encode`` :19-88), backed by
``al_grid_encode_forward / _backward`` (csrc/encoding.cu).

The table is fp32 (the reference keeps fp32 embeddings as well: the autocast down-cast in
grid.py:37-38 is commented out).
"""
import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from ._lib import call, ptr, stream_ptr

_gridtype_to_id = {'hash': 0, 'tiled': 1}


def level_offsets(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size):
    """Entry offset of every level: min(2^log2_hashmap_size, (res+1)^D) rounded up to 8 (grid.py:113-124)."""
    offsets, offset = [], 0
    cap = 2 ** log2_hashmap_size
    for level in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** level))
        n = min(cap, (res + 1) ** input_dim)
        n = int(np.ceil(n / 8) * 8)
        offsets.append(offset)
        offset += n
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32)


class _GridEncode(Function):

    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False,
                gridtype=0):
        if not inputs.is_cuda:
            raise RuntimeError("grid_encode needs CUDA tensors; there is no CPU fallback")
        inputs = inputs.float().contiguous()
        emb = embeddings.float().contiguous()
        B, D = inputs.shape
        L = offsets.shape[0] - 1
        C = emb.shape[1]
        S = float(np.log2(per_level_scale))
        H = int(base_resolution)
        dev = inputs.device
        out = torch.empty(L, B, C, dtype=torch.float32, device=dev)
        dy_dx = torch.empty(B, L * D * C, dtype=torch.float32, device=dev) if calc_grad_inputs else None
        call("al_grid_encode_forward", ptr(inputs), ptr(emb), ptr(offsets), ptr(out), B, D, C, L, S, H,
             1 if calc_grad_inputs else 0, ptr(dy_dx), int(gridtype), None, stream_ptr(dev))
        ctx.save_for_backward(inputs, offsets, dy_dx if dy_dx is not None else torch.empty(0, device=dev))
        ctx.cfg = (B, D, C, L, S, H, int(gridtype), bool(calc_grad_inputs), tuple(emb.shape))
        return out.permute(1, 0, 2).reshape(B, L * C)

    @staticmethod
    def backward(ctx, grad):
        inputs, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, gridtype, calc, emb_shape = ctx.cfg
        dev = inputs.device
        grad = grad.float().view(B, L, C).permute(1, 0, 2).contiguous()
        g_emb = torch.zeros(emb_shape, dtype=torch.float32, device=dev)
        g_in = torch.zeros(B, D, dtype=torch.float32, device=dev) if calc else None
        call("al_grid_encode_backward", ptr(grad), ptr(inputs), ptr(offsets), ptr(g_emb), B, D, C, L, S, H,
             1 if calc else 0, ptr(dy_dx) if calc else None, ptr(g_in), gridtype, stream_ptr(dev))
        return g_in, g_emb, None, None, None, None, None


grid_encode = _GridEncode.apply


def grid_corner_indices(inputs, embeddings, offsets, per_level_scale, base_resolution, gridtype=0):
    """Parity probe: table entry index of every interpolation corner, int32 [B, L, 2^D] (-1 = OOB)."""
    inputs = inputs.float().contiguous()
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    C = embeddings.shape[1]
    dev = inputs.device
    out = torch.empty(L, B, C, dtype=torch.float32, device=dev)
    idx = torch.empty(B, L, 2 ** D, dtype=torch.int32, device=dev)
    call("al_grid_encode_forward", ptr(inputs), ptr(embeddings.float().contiguous()), ptr(offsets), ptr(out), B, D,
         C, L, float(np.log2(per_level_scale)), int(base_resolution), 0, None, int(gridtype), ptr(idx),
         stream_ptr(dev))
    return idx, out


class GridEncoder(nn.Module):
    """Drop-in for torch_ngp.gridencoder.GridEncoder (same constructor, attributes and forward)."""

    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, gridtype='hash'):
        super().__init__()
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]
        self.max_params = 2 ** log2_hashmap_size
        offsets = level_offsets(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size)
        self.register_buffer('offsets', torch.from_numpy(offsets))
        self.n_params = int(offsets[-1]) * level_dim
        self.embeddings = nn.Parameter(torch.empty(int(offsets[-1]), level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)

    def __repr__(self):
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> "
                f"{int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))} "
                f"per_level_scale={self.per_level_scale:.4f} params={tuple(self.embeddings.shape)} "
                f"gridtype={self.gridtype}")

    def forward(self, inputs, bound=1):
        inputs = (inputs + bound) / (2 * bound)
        prefix = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        out = grid_encode(inputs, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution,
                          inputs.requires_grad, self.gridtype_id)
        return out.view(prefix + [self.output_dim])
