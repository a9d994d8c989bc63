import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement (oracle/libbtoracle.so) — the checker, never the product."""
    from tests import _oracle
    return _oracle.load()


@pytest.fixture(scope="session")
def btg():
    """libbtgpu through its C ABI; fails loudly without a GPU / without the .so."""
    from bayestyper_b200 import capi
    lib = capi.load()
    capi.check(lib.btg_init(0), lib)
    yield lib
    lib.btg_shutdown()
