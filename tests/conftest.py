import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "exactdiagonalization.jl_b200"), os.path.join(ROOT, "oracle"), os.path.dirname(__file__)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(os.path.dirname(__file__), "golden", "reference_tests.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ed():
    """The product's host package; importing it loads libedcuda.so (fails loudly if it is not built)."""
    import edcuda
    return edcuda


@pytest.fixture(scope="session")
def gpu_ed(ed):
    if ed.device_count() < 1:
        pytest.fail("GPU test selected but libedcuda sees no CUDA device (there is no CPU fallback)")
    return ed
