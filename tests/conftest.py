import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure).  Builds its C restatement on first use."""
    import subprocess

    so = os.path.join(ROOT, "oracle", "libcheb_oracle.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    import cheb_oracle

    return cheb_oracle


@pytest.fixture(scope="session")
def fc():
    """The product package (needs the built libfastcheb_cuda.so; no fallback)."""
    import fastcheb

    return fastcheb
