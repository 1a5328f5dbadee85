import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure; parity-unpinned against Julia, see oracle/README.md)."""
    from oracle import loader

    loader.lib()
    return loader


@pytest.fixture(scope="session")
def params_cls():
    from adiabatic_raytracer_b200.params import ArtParams

    return ArtParams
