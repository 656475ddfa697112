import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")
    config.addinivalue_line("markers", "fullsize: a BASELINE config at its full size (tens of seconds of host-side set-up each)")


def _has_gpu():
    try:
        import cg_b200
        return cg_b200.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    has_ref = os.path.isdir(REFERENCE)
    for item in items:
        if "reference" in item.keywords and not has_ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference not mounted"))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def cpu_ref():
    import cpu_ref as m
    m.build()
    return m


@pytest.fixture(scope="session")
def gpu():
    import cg_b200
    if cg_b200.device_count() < 1:
        pytest.fail("no CUDA device visible: the gpu tests have no fallback")
    return cg_b200
