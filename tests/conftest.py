import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "nbodygradient.jl_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


# TRAPPIST-1 elements: the reference fixture test/elements.txt == src/ics/defaults.jl:3-16
# (rows: m, P, t0, ecosw, esinw, I, Omega).  Kept as data; committed copy in tests/golden/elements.txt.
def load_elements():
    return np.loadtxt(os.path.join(ROOT, "tests", "golden", "elements.txt"), delimiter=",")


@pytest.fixture(scope="session")
def elements():
    return load_elements()


@pytest.fixture(scope="session")
def oracle():
    from oracle.binding import Oracle, build
    build()
    return Oracle()


def isapprox_maxabs(a, b, rtol=np.sqrt(np.finfo(float).eps)):
    """Julia isapprox(a, b; norm=maxabs) with default rtol = sqrt(eps)."""
    a = np.asarray(a, dtype=float); b = np.asarray(b, dtype=float)
    return np.max(np.abs(a - b)) <= rtol * max(np.max(np.abs(a)), np.max(np.abs(b)))


def tilt(x, v):
    """perturb!() of test/test_integrator.jl:17-25 (x, v are (n,3) [body, dim])."""
    x = x.copy(); v = v.copy()
    x[0, 1] = 5e-1 * np.sqrt(x[0, 0] ** 2 + x[0, 2] ** 2)
    x[1, 1] = -5e-1 * np.sqrt(x[1, 0] ** 2 + x[1, 2] ** 2)
    x[2, 1] = -5e-1 * np.sqrt(x[1, 0] ** 2 + x[1, 2] ** 2)
    v[0, 1] = 5e-1 * np.sqrt(v[0, 0] ** 2 + v[0, 2] ** 2)
    v[1, 1] = -5e-1 * np.sqrt(v[1, 0] ** 2 + v[1, 2] ** 2)
    v[2, 1] = -5e-1 * np.sqrt(v[1, 0] ** 2 + v[1, 2] ** 2)
    return x, v
