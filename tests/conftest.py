import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with `pytest -m gpu`)")


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    d = np.abs(b).max()
    return float(np.abs(a - b).max() / (d if d > 0 else 1.0))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def jf():
    import juliafem.jl_b200 as J
    return J


def curved_tet10(mesh_mod, cx=4, cy=3, cz=3, amp=0.02, seed=7):
    """Kuhn Tet10 box with randomly displaced nodes (curved, non-affine elements)."""
    m = mesh_mod.tet10_kuhn(cx, cy, cz, 1.0)
    h = 1.0 / (2 * cx)
    rng = np.random.default_rng(seed)
    m.coords = m.coords + amp * h * rng.standard_normal(m.coords.shape)
    return m


def distorted_hex8(mesh_mod, n=6, amp=0.1, seed=3):
    m = mesh_mod.hex8_lattice(n, n, n, 1.0 / (n - 1))
    rng = np.random.default_rng(seed)
    m.coords = m.coords + amp / (n - 1) * rng.standard_normal(m.coords.shape)
    return m
