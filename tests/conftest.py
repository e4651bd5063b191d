import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_artifacts():
    """Build the oracle (gcc) and, if missing, the CUDA library (nvcc cross-compiles without a GPU)."""
    from oracle import oracle
    oracle.build()
    from damavand_b200 import build as b
    if not os.path.exists(b.LIB) and os.path.exists(b.nvcc_path()):
        b.build()
    yield
