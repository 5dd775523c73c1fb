"""Pin the CPU oracle (oracle/carma_oracle.cpp) against the reference.

Golden vectors come from the reference's own numpy code run in the build container
(tests/golden/make_golden.py) and from the known answers in the reference's C++ tests.
"""
import numpy as np
import pytest

from oracle import oracle as O


def rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def test_sigsqr_and_variance_known_answer(kelly):
    # carma_unit_tests.cpp:1313-1316  Variance(omega, beta(kappa=.7), sigma=2.3) == 223003.230567 (1e-8 rel)
    from math import comb
    ma07 = np.array([comb(4, i) / 0.7 ** i for i in range(5)])
    v = O.variance(kelly["roots"], ma07, sigma=2.3)
    assert rel(v, 223003.230567) < 1e-8
    assert rel(v, float(kelly["variance_kappa07"])) < 1e-12
    # sigma^2 of the KalmanFilterp/Filter test: 1.1883643935928e-4
    s2 = 2.3 ** 2 / O.variance(kelly["roots"], kelly["ma"], sigma=1.0)
    assert rel(s2, 1.1883643935928e-4) < 1e-12
    assert rel(s2, float(kelly["sigsqr"])) < 1e-13


def test_autocovariance_lags(kelly):
    for lag, want in zip(kelly["acov_lags"], kelly["acov"]):
        got = O.variance(kelly["roots"], kelly["ma"], sigma=np.sqrt(float(kelly["sigsqr"])), lag=float(lag))
        assert rel(got, want) < 1e-11


def test_filterp_matches_reference_numpy(kelly):
    mean, var = O.filterp(kelly["t"], kelly["y"], kelly["yerr"], float(kelly["sigsqr"]), kelly["roots"], kelly["ma"])
    # carma_unit_tests.cpp:441-444: mean(0)==0, var(0)==sigma_y^2+yerr0^2 to 1e-10 abs
    assert mean[0] == 0.0
    assert abs(var[0] - (2.3 ** 2 + kelly["yerr"][0] ** 2)) < 1e-10
    np.testing.assert_allclose(mean, kelly["mean"], rtol=0, atol=2e-10)
    np.testing.assert_allclose(var, kelly["var"], rtol=1e-10, atol=0)
    y = kelly["y"]
    ll = np.cumsum(-0.5 * np.log(var) - 0.5 * (y - mean) ** 2 / var)
    # SURVEY 8(c) golden numbers
    assert abs(ll[269] - 7.828450879851) < 1e-9
    assert abs(ll[499] - 7.909354921621) < 1e-9
    assert abs(ll[999] - 49.705283285319) < 1e-9
    assert abs(ll[999] - float(kelly["loglik_1000"])) < 1e-9 * 50


def test_predict_matches_reference_numpy(kelly):
    qm, qv = O.predictp(kelly["t"], kelly["y"], kelly["yerr"], float(kelly["sigsqr"]), kelly["roots"], kelly["ma"],
                        kelly["predict_t"])
    np.testing.assert_allclose(qm, kelly["predict_mean"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(qv, kelly["predict_var"], rtol=1e-9)
    # interp t=166.0 -> (-1.77608270158, 2.86853702341); forecast -> (-0.00135633527442, 5.28997529984)
    assert abs(qm[0] + 1.77608270158) < 1e-9 and abs(qv[0] - 2.86853702341) < 1e-9
    assert abs(qm[1] + 0.00135633527442) < 1e-11 and abs(qv[1] - 5.28997529984) < 1e-9


def test_predict_vs_bruteforce_gp(kelly):
    """carma_unit_tests.cpp:505-648: Predict == dense-GP conditional with Variance() as kernel (1e-6 rel),
    including backcasting (which the numpy reference does not implement)."""
    n = 120
    t, y, yerr = kelly["t"][:n], kelly["y"][:n], kelly["yerr"][:n]
    s2, roots, ma = float(kelly["sigsqr"]), kelly["roots"], kelly["ma"]
    sig = np.sqrt(s2)
    span = t[-1] - t[0]
    tq = np.array([t[0] - 0.01 * span, 0.5 * (t[40] + t[41]), t[-1] + 0.05 * span])
    qm, qv = O.predictp(t, y, yerr, s2, roots, ma, tq)

    def kern(lag):
        return O.variance(roots, ma, sigma=sig, lag=abs(float(lag)))

    K = np.array([[kern(a - b) for b in t] for a in t]) + np.diag(yerr ** 2)
    for k, tt in enumerate(tq):
        kv = np.array([kern(tt - a) for a in t])
        w = np.linalg.solve(K, kv)
        m = w @ y
        v = kern(0.0) - w @ kv
        assert rel(qm[k], m) < 1e-6
        assert rel(qv[k], v) < 1e-6


def test_scaled_logpost_via_filter(kelly):
    scale, mu = 1.3, 0.25
    mean, var = O.filterp(kelly["t"], kelly["y"] - mu, np.sqrt(scale) * kelly["yerr"], float(kelly["sigsqr"]),
                          kelly["roots"], kelly["ma"])
    ll = np.sum(-0.5 * np.log(var) - 0.5 * (kelly["y"] - mu - mean) ** 2 / var)
    assert abs(ll - 41.593403983566) < 1e-9
    assert abs(ll - float(kelly["scaled_loglik"])) < 1e-9
    assert abs(float(kelly["scaled_logprior"]) + 26.052240106924) < 1e-10


def _series_for(cases, name):
    if name == "c53":
        return cases["t270"], cases["y270"], cases["ysig270"]
    return cases["t60"], cases["y60"], cases["ysig60"]


def test_logdensity_all_pq(loglik_cases):
    from conftest import golden_case_names
    names = golden_case_names(loglik_cases)
    assert len(names) >= 12
    for name in names:
        p, q = int(loglik_cases[name + "_p"]), int(loglik_cases[name + "_q"])
        t, y, e = _series_for(loglik_cases, name)
        kind = O.KIND_CARMA if q > 0 else O.KIND_CARP
        got = O.logdensity(kind, p, q, t, y, e, loglik_cases[name + "_theta"], ignore_prior=True)
        want = loglik_cases[name + "_logpost"]
        assert np.all(np.isfinite(want))
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9, err_msg=name)
        # log prior term (carpack.hpp:118-126)
        for th, lp in zip(loglik_cases[name + "_theta"], loglik_cases[name + "_logprior"]):
            pr = O.default_prior(t, y)
            assert abs(O.log_prior(kind, th, p, pr) - lp) < 1e-12


def test_logdensity_zcarma(loglik_cases):
    t, y, e = loglik_cases["t60"], loglik_cases["y60"], loglik_cases["ysig60"]
    pr = O.default_prior(t, y)
    klo, khi = loglik_cases["z5_kappa_bounds"]
    assert abs(pr.kappa_low - klo) < 1e-15 and abs(pr.kappa_high - khi) < 1e-15
    got = O.logdensity(O.KIND_ZCARMA, 5, 0, t, y, e, loglik_cases["z5_theta"], prior=pr, ignore_prior=True)
    np.testing.assert_allclose(got, loglik_cases["z5_logpost"], rtol=1e-9, atol=1e-9)


def test_zcar_equals_carp(loglik_cases):
    """ZCAR's LogDensity equals plain CAR(p) (shadowed ma_coefs_)."""
    t, y, e = loglik_cases["t60"], loglik_cases["y60"], loglik_cases["ysig60"]
    th = loglik_cases["c50_theta"]
    a = O.logdensity(O.KIND_ZCAR, 5, 0, t, y, e, th, ignore_prior=True)
    b = O.logdensity(O.KIND_CARP, 5, 0, t, y, e, th, ignore_prior=True)
    assert np.array_equal(a, b)


def test_car1_vs_bruteforce_gp(car1_cases):
    t, y, e = car1_cases["t"], car1_cases["y"], car1_cases["yerr"]
    pr = O.default_prior(t, y)
    got = O.logdensity(O.KIND_CAR1, 1, 0, t, y, e, car1_cases["theta"], prior=pr)
    np.testing.assert_allclose(got, car1_cases["logpost"], rtol=1e-9, atol=1e-9)
    th = car1_cases["theta"][0]
    omega = np.exp(th[3])
    mean, var = O.filter1(t, y - th[2], np.sqrt(th[1]) * e, 2 * th[0] ** 2 * omega, omega)
    np.testing.assert_allclose(var, car1_cases["var"][0], rtol=1e-9)
    np.testing.assert_allclose(mean, car1_cases["mean"][0], rtol=0, atol=1e-9)
    # p=1 through the general complex filter is the same model
    m2, v2 = O.filterp(t, y - th[2], np.sqrt(th[1]) * e, 2 * th[0] ** 2 * omega, [-omega + 0j], [1.0])
    np.testing.assert_allclose(v2, var, rtol=1e-10)


def test_prior_bounds(loglik_cases):
    """carma_unit_tests.cpp:1116-1265: each violated bound gives exactly -inf."""
    t, y, e = loglik_cases["t270"], loglik_cases["y270"], loglik_cases["ysig270"]
    pr = O.default_prior(t, y)
    th0 = loglik_cases["c53_theta"][0].copy()
    assert O.check_prior(O.KIND_CARMA, th0, 5, pr)
    assert np.isfinite(O.logdensity(O.KIND_CARMA, 5, 3, t, y, e, th0, prior=pr)[0])

    def bad(th):
        v = O.logdensity(O.KIND_CARMA, 5, 3, t, y, e, th, prior=pr)[0]
        return v == -np.inf

    th = th0.copy(); th[0] = pr.max_stdev * 1.01; assert bad(th)
    th = th0.copy(); th[0] = -0.1; assert bad(th)
    th = th0.copy(); th[1] = 0.49; assert bad(th)
    th = th0.copy(); th[1] = 2.01; assert bad(th)
    # width above max_freq: quad_term2 = -2 Re(w) > 4 pi max_freq
    th = th0.copy(); th[4] = np.log(4 * np.pi * pr.max_freq * 1.5); th[3] = 2 * th[4]; assert bad(th)
    # width below min_freq on the odd real root
    th = th0.copy(); th[7] = np.log(2 * np.pi * pr.min_freq * 0.5); assert bad(th)
    # centroid ordering violated: swap the two pairs
    th = th0.copy(); th[3:5], th[5:7] = th0[5:7].copy(), th0[3:5].copy(); assert bad(th)
    # duplicate roots: identical pairs
    th = th0.copy(); th[5:7] = th[3:5]; assert bad(th)
    # but ignoring the prior evaluates (SetMLE(True), carpack.hpp:180)
    th = th0.copy(); th[1] = 2.01
    assert np.isfinite(O.logdensity(O.KIND_CARMA, 5, 3, t, y, e, th, prior=pr, ignore_prior=True)[0])


def test_long_double_noise_floor(loglik_cases):
    t, y, e = loglik_cases["t270"], loglik_cases["y270"], loglik_cases["ysig270"]
    th = loglik_cases["c53_theta"]
    a = O.logdensity(O.KIND_CARMA, 5, 3, t, y, e, th, ignore_prior=True)
    b = O.logdensity(O.KIND_CARMA, 5, 3, t, y, e, th, ignore_prior=True, long_double=True)
    assert np.max(rel(a, b)) < 1e-10


def test_fast_build_agrees(loglik_cases):
    t, y, e = loglik_cases["t270"], loglik_cases["y270"], loglik_cases["ysig270"]
    th = loglik_cases["c53_theta"]
    a = O.logdensity(O.KIND_CARMA, 5, 3, t, y, e, th, ignore_prior=True)
    b = O.logdensity(O.KIND_CARMA, 5, 3, t, y, e, th, ignore_prior=True, fast=True)
    np.testing.assert_allclose(a, b, rtol=1e-10)
