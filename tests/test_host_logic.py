"""Host-side logic that needs no GPU: synthetic-data helpers against the oracle, the oracle's
MCMC building blocks (RAM adaptation, Cholesky rank-1 update, Philox-driven sampler)."""
import numpy as np
import pytest

from carma_pack_b200 import synth
from oracle import oracle as O


def test_get_ar_roots_and_variance_match_oracle(kelly):
    roots = synth.get_ar_roots(np.array([0.01, 0.01, 0.002]), np.array([0.2, 0.02]))
    assert np.allclose(np.sort_complex(roots), np.sort_complex(kelly["roots"]))
    for lag in (0.0, 2.5):
        a = synth.carma_variance(float(kelly["sigsqr"]), kelly["roots"], kelly["ma"], lag=lag)
        b = O.variance(kelly["roots"], kelly["ma"], sigma=np.sqrt(float(kelly["sigsqr"])), lag=lag)
        assert abs(a - b) < 1e-11 * abs(b)
    assert abs(synth.carma_variance(2.3 ** 2, kelly["roots"], [1, 4 / .7, 6 / .7 ** 2, 4 / .7 ** 3, 1 / .7 ** 4]) -
               223003.230567) < 1e-8 * 223003.230567  # carma_unit_tests.cpp:1313-1316


def test_roots_to_logquad_roundtrip():
    roots, ma, s2 = synth.readme_truth()
    back = O.ar_roots(synth.roots_to_logquad(roots))
    assert np.allclose(np.sort_complex(back), np.sort_complex(roots), rtol=1e-13)
    th = synth.readme_theta(3)
    t, y, e = synth.readme_series(270, 270)
    pr = O.default_prior(t, y)
    assert O.check_prior(O.KIND_CARMA, th, 5, pr)
    beta = O.ma_coefs(O.KIND_CARMA, th, 5, 3, pr)
    assert np.allclose(beta[:3], [1.0, 4.52, 1.34], atol=1e-12) and abs(beta[3] - 0.025) < 1e-12
    assert np.isfinite(O.logdensity(O.KIND_CARMA, 5, 3, t, y, e, th, prior=pr)[0])


def test_carma_process_has_model_variance():
    rng = np.random.default_rng(0)
    roots, ma, s2 = synth.readme_truth()
    t = np.cumsum(rng.uniform(1, 3, 4000))
    y = synth.carma_process(t, s2, roots, ma, rng)
    assert abs(y.std() - 2.3) < 0.35
    # lag-1 autocovariance agrees with the analytic kernel at the mean spacing, roughly
    t2 = np.arange(3000.0) * 2.0
    y2 = synth.carma_process(t2, s2, roots, ma, np.random.default_rng(1))
    emp = np.mean(y2[1:] * y2[:-1])
    assert abs(emp - synth.carma_variance(s2, roots, ma, lag=2.0)) < 1.0


def test_prior_draws_are_mostly_inside_prior():
    t, y, e = synth.readme_series(270, 270)
    rng = np.random.default_rng(3)
    th = synth.prior_draws(400, 5, 3, t, y, rng)
    lp = O.logdensity(O.KIND_CARMA, 5, 3, t, y, e, th)
    assert np.isfinite(lp).mean() > 0.7


def test_chol_update_matches_dense():
    """CholUpdateR1 (steps.cpp:111-131): R'^T R' = R^T R +/- v v^T."""
    rng = np.random.default_rng(2)
    d = 6
    A = rng.standard_normal((d, d))
    S = A @ A.T + d * np.eye(d)
    R = np.linalg.cholesky(S).T
    v = 0.3 * rng.standard_normal(d)
    for down in (False, True):
        R2, _ = O.chol_update(R, v, down)
        want = S + (-1 if down else 1) * np.outer(v, v)
        assert np.allclose(R2.T @ R2, want, rtol=1e-12, atol=1e-12)
        assert np.allclose(np.tril(R2, -1), 0)


def test_tdist_moments():
    x = np.array([O.tdist(11, 3, it, j) for it in range(400) for j in range(11)])
    assert abs(x.mean()) < 0.08
    assert abs(x.var() - 8.0 / 6.0) < 0.25  # var of t_8


def test_oracle_pt_run_consistency():
    """carma_unit_tests.cpp:847-911, 1068-1113: stored log-posteriors equal LogDensity(sample);
    the trace is self-consistent; RAM coerces the acceptance rate toward 0.25."""
    t, y, e = synth.readme_series(90, 5)
    pr = O.default_prior(t, y)
    res = O.pt_run(O.KIND_CARMA, 5, 3, t, y, e, 150, 300, ntemps=4, seed=21, prior=pr, want_trace=True)
    lp = O.logdensity(O.KIND_CARMA, 5, 3, t, y, e, res["samples"], prior=pr)
    np.testing.assert_allclose(lp, res["logposts"], rtol=1e-10)
    rt = res["ram_trace"]
    acc = rt["accepted"].astype(bool)
    drawn = np.isfinite(rt["u"])
    assert np.array_equal(acc[drawn], rt["u"][drawn] < rt["alpha"][drawn])
    assert not acc[~drawn].any() and np.all(rt["alpha"][~drawn] == 0.0)
    assert 0.1 < res["accept_rates"][0] < 0.5
    # a second ensemble index gives a different, equally valid chain; same index reproduces bit for bit
    res2 = O.pt_run(O.KIND_CARMA, 5, 3, t, y, e, 150, 300, ntemps=4, seed=21, prior=pr, ensemble=1)
    res3 = O.pt_run(O.KIND_CARMA, 5, 3, t, y, e, 150, 300, ntemps=4, seed=21, prior=pr)
    assert not np.array_equal(res2["samples"], res["samples"])
    assert np.array_equal(res3["samples"], res["samples"])


def test_oracle_starting_values_finite_and_reproducible():
    t, y, e = synth.readme_series(120, 8)
    pr = O.default_prior(t, y)
    for kind, p, q in [(O.KIND_CARMA, 5, 3), (O.KIND_CARP, 4, 0), (O.KIND_ZCARMA, 3, 0), (O.KIND_CAR1, 1, 0)]:
        th, lp, att = O.starting_value(kind, p, q, t, y, e, pr, seed=5, chain=2)
        th2, lp2, _ = O.starting_value(kind, p, q, t, y, e, pr, seed=5, chain=2)
        assert att >= 0 and np.isfinite(lp) and np.array_equal(th, th2) and lp == lp2
        assert abs(O.logdensity(kind, p, q, t, y, e, th, prior=pr)[0] - lp) < 1e-9 * abs(lp)
