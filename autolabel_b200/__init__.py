"""autolabel_b200 — B200-native (sm_100a) implementation of autolabel's volumetric feature-field
hot path behind the reference's own Python interfaces.

    autolabel_b200.raymarching   <- torch_ngp/raymarching/raymarching.py
    autolabel_b200.gridencoder   <- torch_ngp/gridencoder/grid.py
    autolabel_b200.tcnn          <- tinycudann (the subset autolabel/models.py uses)
    autolabel_b200.renderer      <- torch_ngp/nerf/renderer.py  (NeRFRenderer)
    autolabel_b200.models        <- autolabel/models.py         (ALNetwork)
    autolabel_b200.trainer       <- autolabel/trainer.py        (SimpleTrainer)
    autolabel_b200.parallel      ray-sharded data-parallel training (gradient all-reduce)

All compute goes through libautolabel_b200.so (include/autolabel_b200.h); importing this package
without the built library raises ImportError — there is no CPU or PyTorch fallback.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is missing)

__version__ = "0.1.0"
