"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/autolabel_b200.h declares, and the Python binding table covers the same set."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "autolabel_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(al_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 30
    lib = ctypes.CDLL(os.path.join(ROOT, "autolabel_b200", "libautolabel_b200.so"))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.al_abi_version.restype = ctypes.c_int
    assert lib.al_abi_version() == 1


def test_binding_table_matches_header():
    from autolabel_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared()


def test_mlp_shapes_of_the_named_configs_are_instantiated():
    from autolabel_b200 import _lib
    # C2: hg+freq 48->128->128->16, colour 32->128->128->16, features 16->64->64->64, semantic 80->64->16
    # C1: freq 64->64->64->16, colour 32->64->64->16
    for shape in [(48, 128, 16, 2), (32, 128, 16, 2), (16, 64, 64, 2), (80, 64, 16, 1), (64, 64, 16, 2), (32, 64, 16, 2)]:
        n = _lib.lib.al_mlp_num_params(*shape)
        i, h, o, nh = shape
        assert n == h * i + (h * h if nh == 2 else 0) + o * h
    assert _lib.lib.al_mlp_num_params(7, 7, 7, 7) == -1


def test_no_cpu_fallback():
    import pytest
    import torch
    from autolabel_b200 import raymarching as rm, tcnn
    if torch.cuda.is_available():
        pytest.skip("checks the behaviour without a GPU")
    with pytest.raises(Exception):
        rm.near_far_from_aabb(torch.zeros(4, 3), torch.ones(4, 3), torch.tensor([-1., -1, -1, 1, 1, 1]))
    with pytest.raises(RuntimeError):
        tcnn.Network(15, 64, {"n_neurons": 64, "n_hidden_layers": 2})(torch.zeros(3, 15))
