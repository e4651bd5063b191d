import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a machine without a CUDA device skips the gpu-marked tests (with the reason) instead of
    failing them one by one; `-m gpu` on the B200 box runs them."""
    def n_devices():
        try:
            from damavand_b200 import _lib
            return _lib.load().dvd_device_count()
        except Exception:
            return 0
    if not any("gpu" in it.keywords for it in items):
        return
    if n_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (or libdamavand_b200.so not built): gpu-marked tests run with -m gpu on a B200 box")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_artifacts():
    """Build the oracle (gcc) and, if missing, the CUDA library (nvcc cross-compiles without a GPU)."""
    from oracle import oracle
    oracle.build()
    from damavand_b200 import build as b
    if not os.path.exists(b.LIB) and os.path.exists(b.nvcc_path()):
        b.build()
    yield
