"""Build recipe for oracle/_ref: the reference's OWN CUDA kernels, compiled unmodified.

TEST INFRASTRUCTURE ONLY.  Nothing under autolabel_b200/ may import this.

The reference sources are compiled *where they lie* under /root/reference (never
copied into this repo) with torch.utils.cpp_extension; the only deviation from the
reference's own JIT recipe (torch_ngp/raymarching/backend.py:6-9,
torch_ngp/gridencoder/backend.py:6-9) is -std=c++17 instead of -std=c++14, which the
torch 2.11 headers require, plus an explicit sm_100a target.  No fast-math, exactly as
in the reference flags.

Outputs go to oracle/_ref/ only (git-ignored, but shipped to the GPU box by gpurun):
    oracle/_ref/ref_raymarching.so   pybind module, 11 functions (bindings.cpp:5-19)
    oracle/_ref/ref_gridencoder.so   pybind module, 2 functions (bindings.cpp:5-6)
    oracle/_ref/ref_shencoder.so     pybind module, 2 functions (shencoder/src/bindings.cpp) — pins the SH-4 values
    oracle/_ref/ref_ffmlp.so         pybind module, 5 functions (ffmlp/src/bindings.cpp) — the in-tree bias-free fp16
                                     fully-fused MLP (ffmlp.cu:331-518), the only executable reference-held pin of a
                                     64-wide MLP forward/backward.  Its CUTLASS dependency (an un-vendored submodule,
                                     ffmlp/dependencies/cutlass is empty) is taken from the header tree vendored in this
                                     image (flashinfer/data/cutlass), header-only.

They can only *run* on a CUDA device, so they are used by the `-m gpu` parity tests
(tier O1 in DESIGN.md) and by tests/golden/make_golden.py, which freezes their outputs
into fixtures that the CPU oracle (oracle/ngp_oracle.c) is pinned against.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("AUTOLABEL_REFERENCE", "/root/reference")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
    "-U__CUDA_NO_HALF2_OPERATORS__",
    "-gencode", "arch=compute_100a,code=sm_100a",
]
C_FLAGS = ["-O3", "-std=c++17"]

MODULES = {
    "ref_raymarching": ["torch_ngp/raymarching/src/raymarching.cu",
                        "torch_ngp/raymarching/src/bindings.cpp"],
    "ref_gridencoder": ["torch_ngp/gridencoder/src/gridencoder.cu",
                        "torch_ngp/gridencoder/src/bindings.cpp"],
    "ref_shencoder": ["torch_ngp/shencoder/src/shencoder.cu",
                      "torch_ngp/shencoder/src/bindings.cpp"],
    "ref_ffmlp": ["torch_ngp/ffmlp/src/ffmlp.cu",
                  "torch_ngp/ffmlp/src/bindings.cpp"],
}
# per-module extras (the reference's own flags: torch_ngp/ffmlp/backend.py:6-15)
EXTRA_NVCC = {"ref_ffmlp": ["--expt-extended-lambda", "--expt-relaxed-constexpr", "-Xcompiler=-mf16c",
                            "-Xcompiler=-Wno-float-conversion", "-Xcompiler=-fno-strict-aliasing"]}


def _cutlass_includes():
    """Header-only CUTLASS for ffmlp.cu (its own submodule directory is empty in the reference checkout)."""
    try:
        import flashinfer
        base = os.path.join(os.path.dirname(flashinfer.__file__), "data", "cutlass")
        return [os.path.join(base, "include"), os.path.join(base, "tools", "util", "include")]
    except Exception:
        return []


def available(names=None):
    return all(os.path.exists(os.path.join(OUT, m + ".so")) for m in (names or MODULES))




def build(force=False, verbose=False, only=None):
    """Compile the reference kernels into oracle/_ref. No-op when /root/reference is absent."""
    if not os.path.isdir(REF):
        return False
    os.makedirs(OUT, exist_ok=True)
    from torch.utils.cpp_extension import load
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    for name, srcs in MODULES.items():
        if only and name not in only:
            continue
        target = os.path.join(OUT, name + ".so")
        if os.path.exists(target) and not force:
            continue
        bdir = os.path.join(OUT, "build_" + name)
        os.makedirs(bdir, exist_ok=True)
        load(name=name, sources=[os.path.join(REF, s) for s in srcs],
             extra_cflags=C_FLAGS, extra_cuda_cflags=NVCC_FLAGS + EXTRA_NVCC.get(name, []),
             extra_include_paths=_cutlass_includes() if name == "ref_ffmlp" else None,
             build_directory=bdir, verbose=verbose, is_python_module=False)
        shutil.copy(os.path.join(bdir, name + ".so"), target)
        shutil.rmtree(bdir, ignore_errors=True)
    return True


def load_ref(name):
    """Import a prebuilt reference module (GPU box or here). Raises ImportError if missing."""
    import importlib.util
    import torch  # noqa: F401  (the module links against libtorch)
    path = os.path.join(OUT, name + ".so")
    if not os.path.exists(path):
        raise ImportError(f"{path} not built; run python oracle/build_ref.py where /root/reference exists")
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    only = [a for a in sys.argv[1:] if not a.startswith("-")] or None
    ok = build(force="--force" in sys.argv, verbose=True, only=only)
    print("built" if ok else "reference tree not present; nothing built", OUT)
