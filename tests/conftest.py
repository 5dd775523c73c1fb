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
    # safety net for a checkout whose native libraries were never built (build() of __graft_entry__ normally
    # runs first): compile them once; nvcc cross-compiles sm_100a without a GPU
    if not os.path.exists(os.path.join(ROOT, "carma_pack_b200", "libcarma_b200.so")):
        import build_native
        build_native.build_all(force=False)


@pytest.fixture(scope="session")
def kelly():
    return dict(np.load(os.path.join(GOLDEN, "kelly_carma_test.npz")))


@pytest.fixture(scope="session")
def loglik_cases():
    return dict(np.load(os.path.join(GOLDEN, "loglik_cases.npz")))


@pytest.fixture(scope="session")
def car1_cases():
    return dict(np.load(os.path.join(GOLDEN, "car1_cases.npz")))


def golden_case_names(cases):
    return sorted({k.split("_")[0] for k in cases if k.startswith("c") and k.endswith("_theta")})
