import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def restate():
    """The plain-C restatement of the reference algorithm (oracle/biot_oracle.c) - the checker."""
    from oracle import oracle_py
    return oracle_py.Restatement()


@pytest.fixture(scope="session")
def reference_lib():
    """The reference's own templates behind oracle/ref_driver.cpp; prebuilt oracle/_ref/libo3d_ref.so or skip."""
    from oracle import oracle_py
    if not oracle_py.have_reference():
        pytest.skip("oracle/_ref/libo3d_ref.so not built and /root/reference absent")
    return oracle_py.Reference()


@pytest.fixture(scope="session")
def cuda_ctx():
    """One C-ABI context on cuda:0. Raises (never skips to a fallback) if the library or device is missing."""
    from omega3d_b200.influence import CudaContext
    return CudaContext((0,))


def rel_err(a, b):
    """max |a-b| / max |b| : the norm of SURVEY.md 8d (per-component relative error is meaningless near zeros)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


VEL_TOL = 1e-5   # BASELINE.json north_star: max relative error on velocity
GRAD_TOL = 1e-4  # ... and on the velocity gradient
