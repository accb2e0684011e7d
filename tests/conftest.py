import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ref_rm():
    """The reference's own _raymarching kernels (oracle/_ref, compiled unmodified)."""
    from oracle import build_ref
    try:
        return build_ref.load_ref("ref_raymarching")
    except ImportError as e:
        pytest.skip(str(e))


@pytest.fixture(scope="session")
def ref_ge():
    """The reference's own _gridencoder kernels (oracle/_ref, compiled unmodified)."""
    from oracle import build_ref
    try:
        return build_ref.load_ref("ref_gridencoder")
    except ImportError as e:
        pytest.skip(str(e))
