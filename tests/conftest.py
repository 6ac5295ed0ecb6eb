import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def cref():
    """The literal C oracle (oracle/ref_fftmesh.c), built on demand with gcc."""
    from oracle import cref as m
    m.build()
    return m


@pytest.fixture(scope="session")
def r64():
    from oracle import ref_fft64
    return ref_fft64


@pytest.fixture(scope="session")
def mw():
    """The product package; builds libmistral_ocean.so with nvcc if it is not there yet."""
    lib = os.path.join(ROOT, "mistral-water_b200", "lib", "libmistral_ocean.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()
    import mistral_water_b200
    return mistral_water_b200


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def rel_l2(a, b):
    a = np.asarray(a, np.float64).reshape(-1)
    b = np.asarray(b, np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def max_abs(a, b):
    return float(np.abs(np.asarray(a, np.float64).reshape(-1) - np.asarray(b, np.float64).reshape(-1)).max())
