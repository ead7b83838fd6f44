import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _p8_guard_bands(request):
    """PBT_GUARD=<elements>: every 16-bit activation tensor is allocated between sentinel bands (pbt_b200/_native.py); after each
    GPU test the bands of all live allocations are verified.  Off by default (the bands cost memory and a fill per allocation)."""
    yield
    if os.environ.get("PBT_GUARD", "0") not in ("", "0") and request.node.get_closest_marker("gpu") is not None:
        from pbt_b200._native import check_guards
        check_guards()
