"""ctypes binding of libautolabel_b200.so (the C ABI declared in include/autolabel_b200.h).

There is NO fallback: if the shared library is missing the import fails loudly, and every
wrapper raises RuntimeError (with al_last_error()) on a non-zero return code.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libautolabel_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C autolabel_b200/csrc` (needs nvcc, sm_100a). There is no CPU / PyTorch fallback.")

lib = C.CDLL(LIB_PATH)

P = C.c_void_p
u32 = C.c_uint32
i32 = C.c_int
f32 = C.c_float
f64 = C.c_double
sz = C.c_size_t


class AdamTensor(C.Structure):
    """al_adam_tensor_t"""
    _fields_ = [("param", P), ("grad", P), ("exp_avg", P), ("exp_avg_sq", P), ("n", sz), ("weight_decay", f32)]


class FieldDesc(C.Structure):
    """al_field_t"""
    _fields_ = [
        ("encoding", i32), ("in_pad", i32), ("hidden", i32), ("hidden_color", i32),
        ("feat_dim", i32), ("n_classes", i32), ("bound", f32),
        ("L", u32), ("H", u32), ("gridtype", u32), ("S", f32),
        ("offsets", P), ("table", P), ("w_sigma", P), ("w_color", P), ("w_semf", P), ("w_semo", P),
    ]


_SIGS = {
    "al_last_error": (C.c_char_p, []),
    "al_abi_version": (i32, []),
    "al_sm_count": (i32, []),
    "al_launch_count": (C.c_ulonglong, []),
    "al_near_far_from_aabb": (i32, [P, P, P, u32, f32, P, P, P, P, P]),
    "al_morton3d": (i32, [P, u32, P, P]),
    "al_morton3d_invert": (i32, [P, u32, P, P]),
    "al_packbits": (i32, [P, u32, f32, P, P, P]),
    "al_march_rays_train_workspace": (sz, [u32, u32]),
    "al_march_rays_train": (i32, [P, P, P, f32, f32, u32, u32, u32, u32, u32, P, P, P, f32, P, P, P, P, P,
                                  P, P, P, P, P, P, u32, P, P]),
    "al_march_rays_train_budget": (i32, [P, P, P, f32, f32, u32, u32, u32, u32, u32, P, P, P, P, f32, P, P, P, P, P,
                                         P, P, P, P, P, P, u32, P, P]),
    "al_march_rays_train_count": (i32, [P, P, P, f32, f32, u32, u32, u32, u32, u32, P, P, P, f32, P, P, P, P, P,
                                        u32, P, P]),
    "al_march_rays_train_write": (i32, [P, P, f32, f32, u32, u32, u32, u32, u32, P, P, P, P, P, P, P, P, P]),
    "al_composite_train_fwd": (i32, [P, u32, P, u32, u32, P, P, P, P, u32, u32, f32, P, P, P, P, P, P]),
    "al_composite_train_bwd": (i32, [P, P, P, P, u32, P, u32, u32, P, P, P, P, P, P, u32, u32, f32, P, u32,
                                     P, u32, P, P]),
    "al_composite_train_bwd_weights": (i32, [P, P, P, P, u32, P, u32, u32, P, P, P, P, P, P, u32, u32, f32, P, P, P, P]),
    "al_march_rays": (i32, [u32, u32, P, P, P, P, f32, f32, u32, u32, u32, P, P, P, P, P, P, P, P, u32, P]),
    "al_composite_rays": (i32, [u32, u32, P, P, P, u32, P, u32, u32, P, P, P, f32, P, P, P, P, P, P]),
    "al_compact_rays": (i32, [u32, P, P, P, P, P, P]),
    "al_composite_rays_weights": (i32, [u32, u32, P, P, P, u32, P, P, P, f32, P, P, P, P, P, P]),
    "al_grid_encode_forward": (i32, [P, P, P, P, u32, u32, u32, u32, f32, u32, i32, P, u32, P, P]),
    "al_grid_encode_backward": (i32, [P, P, P, P, u32, u32, u32, u32, f32, u32, i32, P, P, u32, P]),
    "al_freq_encode": (i32, [P, u32, u32, u32, P, P]),
    "al_sh_encode": (i32, [P, u32, P, P]),
    "al_mlp_num_params": (i32, [i32, i32, i32, i32]),
    "al_mlp_forward": (i32, [i32, i32, i32, i32, P, P, i32, i32, P,
                             P, i32, i32, i32, i32, i32,
                             P, i32, i32, i32, i32, i32,
                             P, i32, i32, i32, i32, i32, P]),
    "al_mlp_backward": (i32, [i32, i32, i32, i32, P, P, i32, i32, P, P, i32, i32, i32, P, P, P, i32, i32,
                              i32, i32, P]),
    "al_set_mlp_backend": (i32, [i32]),
    "al_set_bwd_debug": (i32, [i32]),
    "al_mlp_wide_num_params": (i32, [i32, i32, i32, i32]),
    "al_mlp_wide_workspace": (sz, [i32, i32, i32, i32, i32, i32]),
    "al_mlp_wide_forward": (i32, [i32, i32, i32, i32, P, P, i32, i32, P,
                                  P, i32, i32, i32, i32, i32,
                                  P, i32, i32, i32, i32, i32,
                                  P, i32, i32, i32, i32, i32, P, P]),
    "al_mlp_wide_backward": (i32, [i32, i32, i32, i32, P, P, i32, i32, P, P, i32, i32, i32, P, P, P, i32, i32,
                                   i32, P, P]),
    "al_amax": (i32, [P, i32, i32, i32, i32, P, P, P]),
    "al_encode_position": (i32, [P, u32, P, f32, i32, P, P, u32, f32, u32, u32, P, u32, P]),
    "al_head_inputs": (i32, [P, u32, P, P, P, P, P, P, u32, u32, P]),
    "al_grid_scatter_xyz": (i32, [P, u32, P, u32, P, f32, i32, P, P, u32, f32, u32, u32, P]),
    "al_field_workspace": (sz, [C.POINTER(FieldDesc), u32, i32]),
    "al_field_forward": (i32, [C.POINTER(FieldDesc), P, P, P, u32, P, P, u32, P, i32, P, P]),
    "al_field_backward": (i32, [C.POINTER(FieldDesc), P, u32, P, P, P, P, u32, P, P, P, P, P, P, P]),
    "al_field_backward_rays": (i32, [C.POINTER(FieldDesc), P, u32, P, P, u32, P, P, P, P, P, P, P, P, P, P, P, P]),
    "al_density_grid_update": (i32, [P, P, u32, f32, P, P]),
    "al_mark_untrained_grid": (i32, [P, P, u32, f32, f32, f32, f32, f32, u32, u32, P]),
    "al_loss_fwd_bwd": (i32, [P, P, P, u32, u32, u32, P, P, P, P, P, u32, f32, f32, f32, f32, f32, f32, P, P, P, P, P, P]),
    "al_adam_step": (i32, [P, P, P, P, sz, f32, f32, f32, f32, f32, i32, f32, i32, P]),
    "al_adam_multi": (i32, [C.POINTER(AdamTensor), i32, P, P, f64, f64, f32, f32, i32, P]),
    "al_peer_adam_step": (i32, [P, P, P, P, P, P, sz, sz, sz, i32, i32, f32, f32, f32, f32, f32, i32, f32, P]),
    "al_dataset_sample": (i32, [P, P, P, P, P, P, u32, u32, u32, u32, u32, f64, f64, f64, f64, P, i32, P, P, u32, u32,
                                P, P, P, P, P, P, P, P]),
    "al_field_density_pre": (i32, [C.POINTER(FieldDesc), P, u32, P, P, P, P, P]),
    "al_field_workspace_slots": (i32, [C.POINTER(FieldDesc), u32, i32, P, C.POINTER(P), C.POINTER(P)]),
    "al_field_heads_forward": (i32, [C.POINTER(FieldDesc), P, P, u32, P, P, u32, P, P]),
    "al_field_heads_forward_sum": (i32, [C.POINTER(FieldDesc), P, P, u32, P, P, P, u32, i32, P, P]),
    "al_field_density_inputs": (i32, [C.POINTER(FieldDesc), P, P, P, u32, P, P, P, P]),
    "al_compact_alive": (i32, [P, P, P, u32, u32, f32, f32, P, P, P, P, u32, P, P, P, P, P, P, P, P, P, P, u32, P, P]),
    "al_render_epilogue": (i32, [P, P, u32, P, u32, u32, u32, u32, P, u32, P, P, P, P, P, P, P, P, P]),
}

EXPORTS = tuple(_SIGS)

for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)  # AttributeError here = header / library mismatch
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return lib.al_last_error().decode("utf-8", "replace")


def check(code: int, what: str = ""):
    if code != 0:
        raise RuntimeError(f"autolabel_b200: {what} failed with code {code}: {last_error()}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("autolabel_b200 kernels need CUDA tensors; there is no CPU fallback")


def call(name: str, *args):
    check(getattr(lib, name)(*args), name)
