"""N1 / f4: the UNMODIFIED reference `autolabel.models.ALNetwork` constructed over this package's modules through the
`sys.modules` swap INTEGRATION.md (level 2) describes — `tinycudann` -> autolabel_b200.tcnn — and compared with
`autolabel_b200.models.ALNetwork`: same constructor call (autolabel/model_utils.py:61-74), same parameter / buffer names
and shapes, state dicts load into each other, checkpoint keys interchangeable.  Runs where the reference tree exists (the
dev container); the frozen name/shape table (tests/golden/ref_state_dict_spec.json, written by this test module's
`python tests/test_reference_swap_cpu.py`) carries the same check to the GPU box, where /root/reference is absent.
Forward passes need a GPU (no CPU fallback) and are covered by the run()-path goldens (tests/test_renderer_gpu.py)."""
import json
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
REF = os.environ.get("AUTOLABEL_REFERENCE", "/root/reference")
SPEC = os.path.join(ROOT, "tests", "golden", "ref_state_dict_spec.json")

CONFIGS = {
    "c2_hgfreq": dict(encoding='hg+freq', num_layers=2, hidden_dim=128, geo_feat_dim=15, num_layers_color=2, hidden_dim_color=128,
                      hidden_dim_semantic=64, semantic_classes=2, bound=3.0, cuda_ray=True),
    "c1_freq": dict(encoding='freq', num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=2, hidden_dim_color=64,
                    hidden_dim_semantic=64, semantic_classes=2, bound=2.0, cuda_ray=False),
    "c5_lseg": dict(encoding='hg+freq', num_layers=2, hidden_dim=128, geo_feat_dim=15, num_layers_color=2, hidden_dim_color=128,
                    hidden_dim_semantic=512, semantic_classes=2, bound=3.0, cuda_ray=True),
}


def _spec(model):
    sd = model.state_dict()
    return {k: list(v.shape) for k, v in sd.items()}


def _import_reference_over_our_modules():
    """The swap of INTEGRATION.md level 2 (`sys.modules['tinycudann'] = autolabel_b200.tcnn`), with inert stubs for the
    optional third-party imports absent from the image (tests/golden/make_golden_run.py::import_reference)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import autolabel_b200.tcnn as our_tcnn
    from make_golden_run import import_reference
    prev = sys.modules.get("tinycudann")
    try:
        ref_models, _ = import_reference(tcnn_module=our_tcnn)
    finally:
        if prev is not None:
            sys.modules["tinycudann"] = prev
        else:
            sys.modules.pop("tinycudann", None)
    return ref_models


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box): covered by the frozen spec below")
@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_unmodified_reference_model_constructs_over_our_tcnn_and_matches(name):
    ref_models = _import_reference_over_our_modules()
    from autolabel_b200.models import ALNetwork
    kw = CONFIGS[name]
    ref = ref_models.ALNetwork(**kw)
    ours = ALNetwork(**kw)
    a, b = _spec(ref), _spec(ours)
    assert a == b, {k: (a.get(k), b.get(k)) for k in set(a) | set(b) if a.get(k) != b.get(k)}
    # parameters are the tcnn-shaped modules of THIS package on both sides
    assert type(ref.sigma_net).__module__ == "autolabel_b200.tcnn"
    # state dicts load into each other, strictly
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(ours.state_dict(), strict=True)
    for (k1, p1), (k2, p2) in zip(sorted(ref.state_dict().items()), sorted(ours.state_dict().items())):
        assert k1 == k2 and torch.equal(p1, p2)
    # the optimiser groups of scripts/train.py:50-63 address the same parameters
    assert [tuple(p.shape) for p in ref.encoder.parameters()] == [tuple(p.shape) for p in ours.encoder.parameters()]
    assert [tuple(p.shape) for p in ref.network_parameters()] == [tuple(p.shape) for p in ours.network_parameters()]
    # the frozen table (what the GPU box checks) is up to date
    frozen = json.load(open(SPEC))
    assert frozen[name] == a


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_state_dict_matches_frozen_reference_spec(name):
    """Same names and shapes as the reference model's state dict (table frozen from the unmodified reference code)."""
    from autolabel_b200.models import ALNetwork
    frozen = json.load(open(SPEC))
    assert _spec(ALNetwork(**CONFIGS[name])) == frozen[name]


if __name__ == "__main__":
    ref_models = _import_reference_over_our_modules()
    out = {name: _spec(ref_models.ALNetwork(**kw)) for name, kw in CONFIGS.items()}
    json.dump(out, open(SPEC, "w"), indent=1, sort_keys=True)
    print("wrote", SPEC, {k: len(v) for k, v in out.items()})
