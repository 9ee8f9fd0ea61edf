import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def capi():
    from vsr_b200 import _capi

    _capi.build_library()
    return _capi


@pytest.fixture(scope="session")
def real_weights_path():
    p = os.path.join(ROOT, "weights", "sttn-auto", "infer_model.pth")
    if not os.path.exists(p):
        pytest.skip("reference checkpoint not staged under weights/ (tools/stage_weights.py)")
    return p


GOLDEN = os.path.join(ROOT, "tests", "golden")

# The GPU box receives the compact ProPainter checkpoint only (tools/stage_weights.py, snapshot cap): give it the name the tests look for.
_pp = os.path.join(ROOT, "weights", "propainter", "ProPainter.pth")
if not os.path.exists(_pp) and os.path.exists(_pp[:-4] + ".f16.pth"):
    try:
        os.symlink("ProPainter.f16.pth", _pp)
    except OSError:
        pass
