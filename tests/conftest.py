import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def small_inputs():
    """2,000 reads x 5 kb, 30x, e=0.15 -- ~56k output pairs; the oracle needs < 0.1 s."""
    from bella_b200 import frontend as fe
    return fe.synthetic(2000, 5000, seed=7)


@pytest.fixture(scope="session")
def medium_inputs():
    """8,000 reads x 8 kb -- ~0.5 M output pairs."""
    from bella_b200 import frontend as fe
    return fe.synthetic(8000, 8000, seed=11)
