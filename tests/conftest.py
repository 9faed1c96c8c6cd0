import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dune-mmesh_b200", "python"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def gpu_ctx_factory():
    """Creates mmb contexts; fails loudly (no CPU fallback) when the CUDA library or the GPU is missing."""
    from mmesh_b200 import capi
    made = []

    def make(dim=2):
        c = capi.Context(dim=dim, device=0)
        made.append(c)
        return c

    yield make
    for c in made:
        c.close()
