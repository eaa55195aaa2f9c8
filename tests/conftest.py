import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped -- not failed -- on a host without a CUDA device.  Only a missing device skips: a missing or
    broken library still fails (the product has no CPU fallback and the tests must say so)."""
    if not any("gpu" in item.keywords for item in items):
        return
    try:
        from thrifty_b200 import _native
        n_dev = _native.load_library().thr_device_count()
    except Exception:           # noqa: BLE001  library missing / unloadable: let the tests fail loudly
        return
    if n_dev > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device on this host (thr_device_count() == 0)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
