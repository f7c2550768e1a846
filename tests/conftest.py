"""tests/conftest.py — markers and shared fixtures.

`-m "not gpu"`: oracle vs the reference's golden vectors / compiled reference, host logic, ABI surface.
`-m gpu`: parity of the CUDA path (through the C ABI) against the oracle; needs a B200.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Checker libraries (oracle/_ref) and the product library are built once per session if missing."""
    need = [os.path.join(ROOT, "oracle", "_ref", "libhex8_oracle.so"),
            os.path.join(ROOT, "nimblesm_b200", "lib", "libnsm_b200.so"),
            os.path.join(ROOT, "nimblesm_b200", "lib", "libnsm_host_c.so"),
            os.path.join(ROOT, "nimblesm_b200", "lib", "NimbleSM_b200")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__ as g

        g.build()


@pytest.fixture(scope="session")
def oracle():
    from oracle import hex8

    return hex8


@pytest.fixture(scope="session")
def refdrive():
    from oracle import refdrive as r

    if not r.available():
        pytest.skip("oracle/_ref/libnimble_ref.so not built (needs /root/reference at build time)")
    return r


def load_golden(name):
    from tests.golden.make_golden import load_case

    return load_case(name)


def perturbed_cube(n, eps, seed=99):
    """Structured n^3 cube with nodal coordinate noise and a random displacement of relative size eps."""
    from nimblesm_b200.mesh import structured_cube

    mesh = structured_cube(n)
    rng = np.random.default_rng(seed)
    h = 1.0 / n
    ref = np.stack([mesh["x"], mesh["y"], mesh["z"]], 1)
    ref = ref + 0.15 * h * (rng.random(ref.shape) - 0.5)
    mesh["x"], mesh["y"], mesh["z"] = (np.ascontiguousarray(ref[:, i]) for i in range(3))
    disp = eps * h * (2.0 * rng.random(ref.shape) - 1.0)
    return mesh, np.ascontiguousarray(ref), disp
