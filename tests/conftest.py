import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.dirname(os.path.abspath(__file__)) not in sys.path:
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # safety net for a checkout whose native libraries were never built (build() of __graft_entry__ normally
    # runs first): compile them once; nvcc cross-compiles sm_100a without a GPU
    if not os.path.exists(os.path.join(ROOT, "carma_pack_b200", "libcarma_b200.so")):
        import build_native
        build_native.build_all(force=False)


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Audit trail of the noise-floor criterion (tests/parity_util.py): printed even under -q."""
    import parity_util
    rep = parity_util.write_report(ROOT)
    if not rep:
        return
    tr = terminalreporter
    tr.write_line("")
    tr.write_line("log-density parity audit (rtol %.0e; rows that needed the oracle-noise-floor criterion):" % rep["rtol"])
    for r in rep["calls"]:
        tr.write_line("  %-46s rows %7d  noise-floor rows %5d (%.4f%%, allowed %.3g%%)  worst err/noise %.2f  max rel err elsewhere %.2e"
                      % (r["what"][:46], r["rows"], r["noise_floor_rows"], 100 * r["noise_floor_frac"], 100 * r["allowed_frac"],
                         r["worst_err_over_noise"], r["max_rel_err_other_rows"]))
    tr.write_line("  total: %d of %d rows" % (rep["total_noise_floor_rows"], rep["total_rows"]))


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
