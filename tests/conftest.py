import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "openfoam-2.2.x_b200"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ctx():
    """One ldub200.Context on cuda:0 for the whole GPU session."""
    import ldub200
    c = ldub200.Context(0)
    yield c
    c.close()
