"""GPU parity tests: the CUDA path (through the C ABI, via carma_pack_b200._lib) against the CPU
oracle and the committed golden vectors produced by the reference's own numpy code.

Tolerance (BASELINE.json north_star): log-densities within 1e-9 relative of the reference CPU
filter; -inf / NaN must fall in the same class.  For parameter vectors where the reference
algorithm itself is ill-conditioned (near-degenerate AR roots: its own double-precision result
moves by more than the tolerance when re-evaluated in long double), the comparison is made against
that intrinsic noise floor instead; such rows are counted and must stay rare.
"""
import numpy as np
import pytest

from conftest import golden_case_names
from parity_util import RTOL, assert_logpost_parity, ulp_shift

pytestmark = pytest.mark.gpu



@pytest.fixture(scope="module")
def C():
    import carma_pack_b200 as c
    if c._lib.device_count() < 1:
        pytest.fail("no CUDA device visible: GPU tests must run on the B200 box")
    return c


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def to_prior(C, opr):
    return C.Prior(opr.max_stdev, opr.max_freq, opr.min_freq, opr.kappa_low, opr.kappa_high, opr.measerr_dof)


# ------------------------------------------------------------------------------------------------
def test_filter_matches_golden_and_oracle(C, O, kelly):
    s = C.Series(kelly["t"], kelly["y"], kelly["yerr"])
    mean, var = s.filter(float(kelly["sigsqr"]), kelly["roots"], kelly["ma"])
    assert mean[0] == 0.0
    assert abs(var[0] - (2.3 ** 2 + kelly["yerr"][0] ** 2)) < 1e-10  # carma_unit_tests.cpp:441-444
    np.testing.assert_allclose(mean, kelly["mean"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(var, kelly["var"], rtol=1e-9)
    om, ov = O.filterp(kelly["t"], kelly["y"], kelly["yerr"], float(kelly["sigsqr"]), kelly["roots"], kelly["ma"])
    np.testing.assert_allclose(mean, om, rtol=0, atol=1e-9)
    np.testing.assert_allclose(var, ov, rtol=1e-9)
    y = kelly["y"]
    ll = np.cumsum(-0.5 * np.log(var) - 0.5 * (y - mean) ** 2 / var)
    assert abs(ll[269] - 7.828450879851) < 1e-8
    assert abs(ll[999] - 49.705283285319) < 5e-8
    # scaled / centred variant
    m2, v2 = s.filter(float(kelly["sigsqr"]), kelly["roots"], kelly["ma"], measerr_scale=1.3, mu=0.25)
    ll2 = np.sum(-0.5 * np.log(v2) - 0.5 * (y - 0.25 - m2) ** 2 / v2)
    assert abs(ll2 - 41.593403983566) < 5e-8
    s.close()


def test_predict_matches_golden_and_oracle(C, O, kelly):
    s = C.Series(kelly["t"], kelly["y"], kelly["yerr"])
    s2, roots, ma = float(kelly["sigsqr"]), kelly["roots"], kelly["ma"]
    qm, qv = s.predict(s2, roots, ma, kelly["predict_t"])
    np.testing.assert_allclose(qm, kelly["predict_mean"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(qv, kelly["predict_var"], rtol=1e-8)
    t = kelly["t"]
    span = t[-1] - t[0]
    tq = np.concatenate([[t[0] - 0.01 * span, t[0] - 1e-6, t[-1] + 0.05 * span, t[5], t[-1]],
                         np.random.default_rng(0).uniform(t[0], t[-1], 40)])
    qm, qv = s.predict(s2, roots, ma, tq)
    om, ov = O.predictp(t, kelly["y"], kelly["yerr"], s2, roots, ma, tq)
    np.testing.assert_allclose(qm, om, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(qv, ov, rtol=1e-8)
    s.close()


def test_kalmanfilter1_semantics(C, O, car1_cases):
    """KalmanFilter1 == general filter with p=1, omega=(-w,0) (kfilter.cpp:19-48)."""
    t, y, e = car1_cases["t"], car1_cases["y"], car1_cases["yerr"]
    s = C.Series(t, y, e)
    th = car1_cases["theta"][0]
    w = np.exp(th[3])
    mean, var = s.filter(2 * th[0] ** 2 * w, [-w + 0j], [1.0], measerr_scale=th[1], mu=th[2])
    np.testing.assert_allclose(var, car1_cases["var"][0], rtol=1e-9)
    np.testing.assert_allclose(mean, car1_cases["mean"][0], rtol=0, atol=1e-9)
    tq = np.array([t[0] - 3.0, 0.5 * (t[3] + t[4]), t[-1] + 10.0])
    qm, qv = s.predict(2 * th[0] ** 2 * w, [-w + 0j], [1.0], tq, measerr_scale=th[1], mu=th[2])
    om, ov = O.predict1(t, y - th[2], np.sqrt(th[1]) * e, 2 * th[0] ** 2 * w, w, tq)
    np.testing.assert_allclose(qm, om, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(qv, ov, rtol=1e-9)
    s.close()


def _series_for(cases, name):
    if name == "c53":
        return cases["t270"], cases["y270"], cases["ysig270"]
    return cases["t60"], cases["y60"], cases["ysig60"]


def test_loglik_golden_all_pq(C, O, loglik_cases):
    """Every (p,q) family against the reference's numpy LogDensity (golden) and the oracle."""
    for name in golden_case_names(loglik_cases):
        p, q = int(loglik_cases[name + "_p"]), int(loglik_cases[name + "_q"])
        t, y, e = _series_for(loglik_cases, name)
        kind = C.KIND_CARMA if q > 0 else C.KIND_CARP
        s = C.Series(t, y, e)
        th = loglik_cases[name + "_theta"]
        got = s.loglik(kind, p, q, th, flags=C.IGNORE_BOUNDS)
        assert_logpost_parity(got, loglik_cases[name + "_logpost"], what="golden " + name)
        assert_logpost_parity(got, O.logdensity(kind, p, q, t, y, e, th, ignore_prior=True), what="oracle " + name)
        # LOGLIK_ONLY drops exactly the prior term
        got_ll = s.loglik(kind, p, q, th, flags=C.IGNORE_BOUNDS | C.LOGLIK_ONLY)
        np.testing.assert_allclose(got_ll, loglik_cases[name + "_loglik"], rtol=1e-9, atol=1e-9)
        s.close()


def test_loglik_zcarma_zcar_car1(C, O, loglik_cases, car1_cases):
    t, y, e = loglik_cases["t60"], loglik_cases["y60"], loglik_cases["ysig60"]
    s = C.Series(t, y, e)
    pr = s.default_prior()
    opr = O.default_prior(t, y)
    assert np.allclose(pr.as_tuple(), (opr.max_stdev, opr.max_freq, opr.min_freq, opr.kappa_low, opr.kappa_high,
                                       opr.measerr_dof), rtol=1e-14)
    got = s.loglik(C.KIND_ZCARMA, 5, 0, loglik_cases["z5_theta"], prior=pr, flags=C.IGNORE_BOUNDS)
    assert_logpost_parity(got, loglik_cases["z5_logpost"], what="zcarma golden")
    # ZCAR is evaluated as CAR(p) by the reference
    a = s.loglik(C.KIND_ZCAR, 5, 0, loglik_cases["c50_theta"], flags=C.IGNORE_BOUNDS)
    b = s.loglik(C.KIND_CARP, 5, 0, loglik_cases["c50_theta"], flags=C.IGNORE_BOUNDS)
    assert np.array_equal(a, b)
    s.close()
    t, y, e = car1_cases["t"], car1_cases["y"], car1_cases["yerr"]
    s = C.Series(t, y, e)
    got = s.loglik(C.KIND_CAR1, 1, 0, car1_cases["theta"])
    assert_logpost_parity(got, car1_cases["logpost"], what="car1 brute-force GP")
    assert_logpost_parity(got, O.logdensity(O.KIND_CAR1, 1, 0, t, y, e, car1_cases["theta"]), what="car1 oracle")
    s.close()


def test_prior_bounds_give_minus_inf(C, loglik_cases):
    """carma_unit_tests.cpp:1116-1265"""
    t, y, e = loglik_cases["t270"], loglik_cases["y270"], loglik_cases["ysig270"]
    s = C.Series(t, y, e)
    pr = s.default_prior()
    th0 = loglik_cases["c53_theta"][0].copy()
    rows = [th0.copy()]
    th = th0.copy(); th[0] = pr.max_stdev * 1.01; rows.append(th)
    th = th0.copy(); th[0] = -0.1; rows.append(th)
    th = th0.copy(); th[1] = 0.49; rows.append(th)
    th = th0.copy(); th[1] = 2.01; rows.append(th)
    th = th0.copy(); th[4] = np.log(4 * np.pi * pr.max_freq * 1.5); th[3] = 2 * th[4]; rows.append(th)
    th = th0.copy(); th[7] = np.log(2 * np.pi * pr.min_freq * 0.5); rows.append(th)
    th = th0.copy(); th[3:5], th[5:7] = th0[5:7].copy(), th0[3:5].copy(); rows.append(th)
    th = th0.copy(); th[5:7] = th[3:5]; rows.append(th)
    got = s.loglik(C.KIND_CARMA, 5, 3, np.array(rows), prior=pr)
    assert np.isfinite(got[0])
    assert np.all(got[1:] == -np.inf)
    got = s.loglik(C.KIND_CARMA, 5, 3, np.array(rows), prior=pr, flags=C.IGNORE_BOUNDS)
    assert np.isfinite(got[3]) and np.isfinite(got[4])
    s.close()


def test_loglik_config2_65536_thetas(C, O):
    """BASELINE config 2: 65,536 CARMA(5,3) parameter vectors on one ny=270 series, every row against the oracle."""
    from carma_pack_b200 import synth
    t, y, e = synth.readme_series(270, 270)
    th = synth.theta_batch(65536, t, y, seed=2)
    s = C.Series(t, y, e)
    pr = s.default_prior()
    got = s.loglik(C.KIND_CARMA, 5, 3, th, prior=pr)
    opr = O.default_prior(t, y)
    want = O.logdensity(O.KIND_CARMA, 5, 3, t, y, e, th, prior=opr)
    assert 0.5 < np.isfinite(want).mean() < 1.0  # both the filter path and the -inf early-out are exercised
    want_ld = O.logdensity(O.KIND_CARMA, 5, 3, t, y, e, th, prior=opr, long_double=True)
    # audited ceiling: at most 40 of the 65,536 rows may need the noise-floor criterion (round 2 observed: ~20)
    n_ill = assert_logpost_parity(
        got, want, want_ld, what="config2", max_illcond_rows=40,
        ulp_eval=lambda rows, k: O.logdensity(O.KIND_CARMA, 5, 3, t, y, e, ulp_shift(th[rows], k), prior=opr))
    print("config2: %d of 65536 rows used the noise-floor criterion" % n_ill)
    # idempotence / determinism: same input, same bits
    got2 = s.loglik(C.KIND_CARMA, 5, 3, th, prior=pr)
    assert np.array_equal(got, got2, equal_nan=True)
    s.close()


def test_loglik_long_series_chunked_pipeline(C, O, kelly):
    """ny = 1000 and 5000 exercise the double-buffered TMA pipeline (chunks of 512 points)."""
    from carma_pack_b200 import synth
    rng = np.random.default_rng(5)
    for ny in (1000, 5001):
        t, y, e = synth.readme_series(ny, ny)
        th = synth.theta_batch(256, t, y, seed=ny)
        th[1] = th[0]  # duplicates must agree bitwise
        s = C.Series(t, y, e)
        got = s.loglik(C.KIND_CARMA, 5, 3, th)
        want = O.logdensity(O.KIND_CARMA, 5, 3, t, y, e, th)
        want_ld = O.logdensity(O.KIND_CARMA, 5, 3, t, y, e, th, long_double=True)
        assert_logpost_parity(
            got, want, want_ld, max_illcond_frac=0.02, what="ny=%d" % ny,
            ulp_eval=lambda rows, k: O.logdensity(O.KIND_CARMA, 5, 3, t, y, e, ulp_shift(th[rows], k)))
        assert got[0] == got[1] or (np.isnan(got[0]) and np.isnan(got[1]))
        s.close()
    # kelly fixture through LogDensity-like path is covered by the golden filter test; here the
    # time-shift invariance property at full length: only dt enters the filter
    t, y, e = kelly["t"], kelly["y"], kelly["yerr"]
    th = synth.prior_draws(64, 4, 2, t, y, rng)
    s1, s2 = C.Series(t, y, e), C.Series(t + 1000.0, y, e)
    a = s1.loglik(C.KIND_CARMA, 4, 2, th, flags=C.IGNORE_BOUNDS)
    b = s2.loglik(C.KIND_CARMA, 4, 2, th, flags=C.IGNORE_BOUNDS)
    np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-8, equal_nan=True)
    s1.close(); s2.close()


def test_loglik_stress_all_orders(C, O):
    """Every (p,q) of the choose_order grid (p = 1..7, q < p), random parameter vectors including overdamped
    (real-pair) roots and out-of-prior rows, with and without bounds, on an irregular series."""
    from carma_pack_b200 import synth
    rng = np.random.default_rng(2024)
    t = synth.cauchy_times(140, rng)
    y = 1.0 + np.cumsum(rng.standard_normal(140)) * 0.1 + 0.05 * rng.standard_normal(140)
    e = 0.05 * rng.uniform(0.5, 2.0, 140)
    s = C.Series(t, y, e)
    pr = s.default_prior()
    opr = O.default_prior(t, y)
    total_ill = 0
    for p in range(1, 8):
        for q in range(p):
            if p == 1:
                kind, okind = C.KIND_CAR1, O.KIND_CAR1
                th = np.column_stack([y.std() * np.exp(0.3 * rng.standard_normal(150)), rng.uniform(0.4, 2.1, 150),
                                      y.mean() + 0.2 * rng.standard_normal(150), rng.uniform(-8, 3, 150)])
            else:
                kind, okind = (C.KIND_CARMA, O.KIND_CARMA) if q > 0 else (C.KIND_CARP, O.KIND_CARP)
                th = synth.prior_draws(150, p, q, t, y, rng)
                # push a third of the rows into the overdamped regime (two real roots in the first factor)
                th[::3, 3] = th[::3, 4] * 2 - np.log(4.0) - rng.uniform(0.1, 2.0, th[::3].shape[0])
            for flags, ign in ((0, False), (C.IGNORE_BOUNDS, True)):
                if p == 1 and ign:
                    continue
                got = s.loglik(kind, p, q, th, prior=pr, flags=flags)
                want = O.logdensity(okind, p, q, t, y, e, th, prior=opr, ignore_prior=ign)
                want_ld = O.logdensity(okind, p, q, t, y, e, th, prior=opr, ignore_prior=ign, long_double=True)
                total_ill += assert_logpost_parity(
                    got, want, want_ld, max_illcond_frac=0.08, what="stress p=%d q=%d ign=%d" % (p, q, ign),
                    ulp_eval=lambda rows, k: O.logdensity(okind, p, q, t, y, e, ulp_shift(th[rows], k), prior=opr,
                                                          ignore_prior=ign))
    print("stress: %d rows needed the noise-floor criterion" % total_ill)
    assert total_ill <= 120, total_ill   # of 7,950 rows over all orders (observed in round 2: ~45; 4 % in the worst (p,q))
    s.close()


def test_async_pipeline_equals_blocking_call(C):
    """carma_loglik_batch_async/_wait (CARMA_N_SLOTS = 4 slots) returns exactly what the blocking call returns, with
    two and with four steps in flight; the blocking call itself (four pieces on four streams for >= 16,384 rows)
    equals one launch over device-resident rows."""
    import ctypes
    import torch
    from carma_pack_b200 import synth
    t, y, e = synth.readme_series(270, 3)
    s = C.Series(t, y, e)
    pr = s.default_prior()
    batches = [np.ascontiguousarray(synth.theta_batch(4096 + 512 * k, t, y, seed=k)) for k in range(9)]
    want = [s.loglik(C.KIND_CARMA, 5, 3, b, prior=pr) for b in batches]
    for nslot in (2, 4):
        outs = [np.empty(b.shape[0]) for b in batches]
        for k, b in enumerate(batches):
            slot = k % nslot
            if k >= nslot:
                s.loglik_wait(slot)
            s.loglik_async(C.KIND_CARMA, 5, 3, b.ctypes.data, outs[k].ctypes.data, b.shape[0], pr, slot)
        for slot in range(nslot):
            s.loglik_wait(slot)
        for k in range(len(batches)):
            assert np.array_equal(outs[k], want[k], equal_nan=True)
    with pytest.raises(C.CarmaError):
        s.loglik_async(C.KIND_CARMA, 5, 3, batches[0].ctypes.data, outs[0].ctypes.data, 16, pr, 4)
    big = np.ascontiguousarray(synth.theta_batch(40000, t, y, seed=77))
    got = s.loglik(C.KIND_CARMA, 5, 3, big, prior=pr)              # four pieces
    d_th = torch.from_numpy(big).cuda()
    d_out = torch.empty(big.shape[0], dtype=torch.float64, device="cuda")
    s.loglik_dev(C.KIND_CARMA, 5, 3, d_th.data_ptr(), d_out.data_ptr(), big.shape[0], pr, 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(got, d_out.cpu().numpy(), equal_nan=True)
    s.close()


def test_scaling_property(C):
    """Size-independent property: (y, yerr, sigma_y, mu) -> c * (...) shifts the log-likelihood by -ny log c."""
    from carma_pack_b200 import synth
    t, y, e = synth.readme_series(270, 11)
    th = synth.theta_batch(512, t, y, seed=4)
    c = 3.7
    th2 = th.copy(); th2[:, 0] *= c; th2[:, 2] *= c
    s1, s2 = C.Series(t, y, e), C.Series(t, c * y, c * e)
    a = s1.loglik(C.KIND_CARMA, 5, 3, th, flags=C.IGNORE_BOUNDS | C.LOGLIK_ONLY)
    b = s2.loglik(C.KIND_CARMA, 5, 3, th2, flags=C.IGNORE_BOUNDS | C.LOGLIK_ONLY)
    fin = np.isfinite(a) & np.isfinite(b)
    assert fin.mean() > 0.9
    np.testing.assert_allclose(b[fin], a[fin] - 270 * np.log(c), rtol=1e-8, atol=1e-6)
    s1.close(); s2.close()


def test_multi_series_ragged(C, O):
    """K4: ragged batch, one theta per curve, per-curve priors."""
    from carma_pack_b200 import synth
    rng = np.random.default_rng(9)
    ar, ma, s2 = synth.carma31_truth()
    ts, ys, es, off, thetas = [], [], [], [0], []
    ncurves = 300
    for c in range(ncurves):
        ny = int(rng.integers(20, 140)) if c else 2  # include the minimum length
        t = synth.cauchy_times(ny, rng)
        y0 = 5.0 + synth.carma_process(t, s2, ar, ma, rng)
        e = np.full(ny, 0.1) * rng.uniform(0.5, 2.0, ny)
        y = y0 + e * rng.standard_normal(ny)
        ts.append(t); ys.append(y); es.append(e); off.append(off[-1] + ny)
        thetas.append(synth.prior_draws(1, 3, 1, t, y, rng)[0] if ny > 4 else np.array([1.0, 1.0, 5.0, -1.0, -0.5, -3.0, 0.5]))
    t, y, e = np.concatenate(ts), np.concatenate(ys), np.concatenate(es)
    th = np.array(thetas)
    m = C.MultiSeries(t, y, e, off)
    pri = m.default_priors()
    got = m.loglik(C.KIND_CARMA, 3, 1, th)
    opri = [O.default_prior(ts[c], ys[c]) for c in range(ncurves)]
    for c in (0, 1, 17):
        assert np.allclose(tuple(pri[c]), (opri[c].max_stdev, opri[c].max_freq, opri[c].min_freq, opri[c].kappa_low,
                                           opri[c].kappa_high, 50.0), rtol=1e-13)
    want = O.logdensity_multi(O.KIND_CARMA, 3, 1, t, y, e, off, th, opri)
    assert_logpost_parity(got, want, what="multi")
    got2 = m.loglik(C.KIND_CARMA, 3, 1, th, flags=C.IGNORE_BOUNDS)
    want2 = O.logdensity_multi(O.KIND_CARMA, 3, 1, t, y, e, off, th, opri, ignore_prior=True)
    assert_logpost_parity(got2, want2, rtol=1e-8, what="multi ignore-bounds")
    m.close()


def test_multi_series_short_curves_at_every_alignment(C, O):
    """K4 fetches a curve in 32-byte blocks of four steps after peeling a misaligned start: curves of 2..17 points
    starting at every offset modulo 4 (heads of 0..3 scalar steps, zero or more blocks, tails of 0..3 steps)."""
    from carma_pack_b200 import synth
    rng = np.random.default_rng(19)
    ar, ma, s2 = synth.carma31_truth()
    lens = [int(n) for n in rng.permutation(np.repeat(np.arange(2, 18), 4))]
    ts, ys, es, off = [], [], [], [0]
    for ny in lens:
        t = synth.cauchy_times(ny, rng)
        y = 5.0 + synth.carma_process(t, s2, ar, ma, rng) + 0.1 * rng.standard_normal(ny)
        ts.append(t); ys.append(y); es.append(np.full(ny, 0.1) * rng.uniform(0.5, 2.0, ny)); off.append(off[-1] + ny)
    assert {o % 4 for o in off[:-1]} == {0, 1, 2, 3}
    t, y, e = np.concatenate(ts), np.concatenate(ys), np.concatenate(es)
    th = np.tile(np.array([1.0, 1.0, 5.0, -1.0, -0.5, -3.0, 0.5]), (len(lens), 1))
    th[:, 2] += 0.05 * rng.standard_normal(len(lens))
    m = C.MultiSeries(t, y, e, off)
    got = m.loglik(C.KIND_CARMA, 3, 1, th, flags=C.IGNORE_BOUNDS)
    opri = [O.default_prior(ts[c], ys[c]) for c in range(len(lens))]
    want = O.logdensity_multi(O.KIND_CARMA, 3, 1, t, y, e, off, th, opri, ignore_prior=True)
    assert np.isfinite(want).all()
    np.testing.assert_allclose(got, want, rtol=1e-9)
    # and the same curves one by one through K1 (series in shared memory): same recursion, different loads
    for c in (0, 5, 17, 40, len(lens) - 1):
        s1 = C.Series(ts[c], ys[c], es[c])
        one = s1.loglik(C.KIND_CARMA, 3, 1, th[c:c + 1], prior=s1.default_prior(), flags=C.IGNORE_BOUNDS)
        np.testing.assert_allclose(one, got[c:c + 1], rtol=1e-11)
        s1.close()
    m.close()


def test_philox_bit_exact_and_tdist(C, O):
    rng = np.random.default_rng(1)
    for _ in range(20):
        c = [int(x) for x in rng.integers(0, 2 ** 32, 4)]
        seed = int(rng.integers(0, 2 ** 63))
        assert C._lib.philox_dev(c[0], c[1], c[2], c[3], seed) == O.philox(c[0], c[1], c[2], c[3], seed)
    # Random123 known-answer vector for Philox4x32-10: counter = key = 0
    assert O.philox(0, 0, 0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert C._lib.philox_dev(0, 0, 0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    for j in range(30):
        a = C._lib.tdist_dev(77, 5, j, j % 11, 8)
        b = O.tdist(77, 5, j, j % 11, 8)
        assert abs(a - b) <= 1e-13 * max(1.0, abs(b))


def test_scan_kalman_matches_sequential_and_oracle(C, O):
    """K5 (associative-scan Kalman): same log-density as the sequential kernel and the oracle for every
    chunking, model family and root type; 1e-9 relative on the total (hard part 8)."""
    from carma_pack_b200 import synth
    rng = np.random.default_rng(17)
    t, y, e = synth.readme_series(1000, 5)
    s = C.Series(t, y, e)
    pr = s.default_prior()
    opr = O.default_prior(t, y)
    for kind_name, p, q in [("CARMA", 5, 3), ("CARMA", 3, 1), ("CARP", 2, 0), ("CARP", 4, 0), ("CARMA", 7, 2),
                            ("CAR1", 1, 0), ("ZCARMA", 3, 0)]:
        kind, okind = getattr(C, "KIND_" + kind_name), getattr(O, "KIND_" + kind_name)
        if kind_name == "CAR1":
            th = np.array([[y.std(), 1.0, y.mean(), np.log(0.05)], [y.std() * 2, 1.2, y.mean() + 1, np.log(0.3)]])
        else:
            th = synth.prior_draws(6, p, q, t, y, rng)
            if kind_name == "ZCARMA":
                th = np.hstack([th, rng.uniform(-2, 2, (6, 1))])
            th[0, 3], th[0, 4] = np.log(0.02), np.log(0.9)          # overdamped pair: two real roots
            th[1, 0] = -1.0                                          # outside the prior: -inf
        seq = s.loglik(kind, p, q, th, prior=pr)
        want = O.logdensity(okind, p, q, t, y, e, th, prior=opr)
        for chunk in (2, 7, 128, 4096):
            got = s.loglik_scan(kind, p, q, th, prior=pr, chunk=chunk)
            assert np.array_equal(np.isfinite(got), np.isfinite(seq)), (kind_name, chunk)
            fin = np.isfinite(seq)
            assert np.all(got[~fin] == seq[~fin])
            np.testing.assert_allclose(got[fin], seq[fin], rtol=1e-9, err_msg="%s chunk=%d vs sequential" % (kind_name, chunk))
            assert_logpost_parity(got, want, O.logdensity(okind, p, q, t, y, e, th, prior=opr, long_double=True),
                                  max_illcond_frac=0.17, what="scan %s chunk=%d" % (kind_name, chunk),
                                  ulp_eval=lambda rows, k: O.logdensity(okind, p, q, t, y, e, ulp_shift(th[rows], k), prior=opr))
    s.close()
    # a long series: 200,000 points, CARMA(3,1), against the CPU oracle
    ar, ma, s2 = synth.carma31_truth()
    n = 200000
    tl = np.cumsum(rng.uniform(0.5, 1.5, n))
    yl = rng.standard_normal(n)  # the likelihood of any data is a valid test; no need to simulate the process
    el = np.full(n, 0.3)
    sl = C.Series(tl, yl, el)
    th = np.array([[1.0, 1.0, 0.0] + list(synth.roots_to_logquad(ar)) + [np.log(1.0 / 3.0)]])
    got = sl.loglik_scan(C.KIND_CARMA, 3, 1, th, flags=C.IGNORE_BOUNDS)
    seq = sl.loglik(C.KIND_CARMA, 3, 1, th, flags=C.IGNORE_BOUNDS)
    want = O.logdensity(O.KIND_CARMA, 3, 1, tl, yl, el, th, ignore_prior=True)
    assert abs(got[0] - seq[0]) <= 1e-9 * abs(seq[0])
    assert abs(got[0] - want[0]) <= 1e-9 * abs(want[0])
    sl.close()


def test_fast_math_accuracy(C):
    """exp_scaled / rot_scaled / rcp_fast (csrc/fast_math.cuh) on the device against long double / exact rational
    argument reduction: the error budget of one transition factor of the time loop."""
    from fractions import Fraction
    rng = np.random.default_rng(0)
    n = 40000
    lnat = -np.exp(rng.uniform(np.log(1e-6), np.log(50.0), n))
    dt = np.exp(rng.uniform(np.log(1e-2), np.log(30.0), n))
    l_exp = lnat * (64.0 / np.log(2.0))
    ex, _, _, s_r, c_r, _ = C._lib.fastmath_dev(l_exp, dt)
    x = l_exp.astype(np.longdouble) * dt.astype(np.longdouble) * (np.log(np.longdouble(2)) / 64)
    want = np.exp(x)
    ok = x > -700
    rel = np.abs(ex[ok] - want[ok]) / want[ok]
    assert rel.max() < 4e-16, rel.max()
    assert np.all(ex[~ok] < 1e-300) and np.all(ex >= 0)   # clamped exponent instead of a flush branch
    assert np.abs(c_r - (1 + want) / 2).max() < 3e-16 and np.abs(s_r - (1 - want) / 2).max() < 3e-16
    lp = -np.exp(rng.uniform(np.log(1e-3), np.log(1e7), n))
    _, s_c, c_c, _, _, rc = C._lib.fastmath_dev(lp, dt)
    assert not np.isnan(s_c).any()   # NaN marks a bit difference between the all-conjugate and the generic variant
    pick = rng.choice(n, 3000, replace=False)
    red = np.array([float(((Fraction(float(lp[i])) * Fraction(float(dt[i]))) % 128)) for i in pick], dtype=np.longdouble)
    lo = np.array([float((Fraction(float(lp[i])) * Fraction(float(dt[i]))) % 128 - Fraction(float(red[k]))) for k, i in enumerate(pick)],
                  dtype=np.longdouble)
    ang = (red + lo) * np.longdouble(np.pi) / 64 + (red + lo) * np.longdouble(1.2246467991473532e-16) / 64
    assert np.abs(s_c[pick] - np.sin(ang)).max() < 4e-16 and np.abs(c_c[pick] - np.cos(ang)).max() < 4e-16
    xr = np.concatenate([np.exp(rng.uniform(np.log(1e-200), np.log(1e200), 20000)), -np.exp(rng.uniform(-5, 5, 100))])
    r = C._lib.fastmath_dev(xr, np.ones_like(xr))[5]
    assert (np.abs(r * xr - 1.0)).max() < 5e-16


# ------------------------------------------------------------------------------------------------
# PT-MCMC
# ------------------------------------------------------------------------------------------------
def _check_trace_against_oracle(C, O, res, kind, p, q, t, y, e, opr, ntemps, tmax=100.0):
    """Record/replay: every proposal's log-density, and every accept / exchange decision, re-derived
    on the CPU from the recorded state."""
    temps = np.exp(np.linspace(0.0, np.log(tmax), ntemps)) if ntemps > 1 else np.ones(1)
    rt, xt, prop = res["ram_trace"][0], res["exchange_trace"][0], res["proposals"][0]
    iters = rt.shape[0]
    d = prop.shape[-1]
    lp_o = O.logdensity(kind, p, q, t, y, e, prop.reshape(-1, d), prior=opr).reshape(iters, ntemps)
    lp_ld = O.logdensity(kind, p, q, t, y, e, prop.reshape(-1, d), prior=opr, long_double=True).reshape(iters, ntemps)
    flat = prop.reshape(-1, d)
    assert_logpost_parity(rt["lp_prop"].ravel(), lp_o.ravel(), lp_ld.ravel(), max_illcond_frac=0.01, what="pt proposals",
                          ulp_eval=lambda rows, k: O.logdensity(kind, p, q, t, y, e, ulp_shift(flat[rows], k), prior=opr))
    nflip = 0
    for it in range(iters):
        for c in range(ntemps):
            r = rt[it, c]
            a = (lp_o[it, c] - r["lp_cur"]) / temps[c]
            if not np.isfinite(a):
                assert r["accepted"] == 0 and r["alpha"] == 0.0
                continue
            alpha = min(np.exp(a), 1.0)
            # d(alpha)/alpha = d(lp)/T: the parity band of the two log-densities (1e-9 relative, or the oracle's
            # own double-vs-long-double noise on an ill-conditioned proposal, as in assert_logpost_parity), floor 1e-7
            noise = abs(lp_o[it, c] - lp_ld[it, c]) if np.isfinite(lp_ld[it, c]) else 0.0
            tol = max(1e-7, (2e-9 * (abs(lp_o[it, c]) + abs(r["lp_cur"])) + 50.0 * noise) / temps[c])
            assert abs(alpha - r["alpha"]) <= tol * max(alpha, 1e-30) + 1e-300
            want_acc = r["u"] < alpha
            if want_acc != bool(r["accepted"]):
                # only allowed when u sits within the tolerance band of alpha
                assert abs(r["u"] - alpha) <= tol * alpha
                nflip += 1
            if c > 0:
                x = xt[it, c]
                ax = (x["lp_prop"] - x["lp_cur"]) / temps[c] + (x["lp_cur"] - x["lp_prop"]) / temps[c - 1]
                with np.errstate(over="ignore"):
                    ax = min(np.exp(ax), 1.0)
                if not np.isfinite(ax):
                    ax = 0.0
                assert abs(ax - x["alpha"]) <= 1e-12 * max(ax, 1e-30) + 1e-300
                assert bool(x["accepted"]) == (x["u"] < x["alpha"])
    return nflip


@pytest.mark.parametrize("kind_name,p,q,ntemps", [("CARMA", 5, 3, 10), ("CARP", 3, 0, 4), ("CAR1", 1, 0, 1),
                                                  ("ZCARMA", 4, 0, 3), ("CARMA", 7, 4, 6), ("CARMA", 2, 1, 2)])
def test_pt_run_record_replay_and_trajectory(C, O, kind_name, p, q, ntemps):
    """The PT kernel against the oracle: every recorded proposal, every decision, and the whole trajectory."""
    from carma_pack_b200 import synth
    kind = getattr(C, "KIND_" + kind_name)
    t, y, e = synth.readme_series(120, 42)
    s = C.Series(t, y, e)
    pr = s.default_prior()
    opr = O.default_prior(t, y)
    nsamples, burnin, thin = 60, 80, 2
    res = s.pt_run(kind, p, q, nsamples, burnin, thin=thin, ntemps=ntemps, n_ensembles=3, seed=1234, prior=pr,
                   record_trace=True)
    assert np.all(np.isfinite(res["logposts"]))
    # stored log-posteriors equal a fresh LogDensity of the stored samples (carma_unit_tests.cpp:1068-1113, 1e-8 rel)
    d = res["samples"].shape[-1]
    relp = s.loglik(kind, p, q, res["samples"].reshape(-1, d), prior=pr).reshape(res["logposts"].shape)
    np.testing.assert_allclose(relp, res["logposts"], rtol=1e-8)
    nflip = _check_trace_against_oracle(C, O, res, kind, p, q, t, y, e, opr, ntemps)
    assert nflip == 0
    # identical Philox streams => the oracle's sequential-order sampler follows the same trajectory
    ores = O.pt_run(kind, p, q, t, y, e, nsamples, burnin, thin=thin, ntemps=ntemps, seed=1234, ensemble=0, prior=opr,
                    want_trace=True)
    acc_g = res["ram_trace"][0]["accepted"]
    acc_o = ores["ram_trace"]["accepted"]
    x_g = res["exchange_trace"][0]["accepted"][:, 1:]
    x_o = ores["exchange_trace"]["accepted"][:, 1:]
    assert np.array_equal(acc_g, acc_o), "accept decisions diverged at iteration %d" % np.argmax((acc_g != acc_o).any(axis=1))
    assert np.array_equal(x_g, x_o)
    np.testing.assert_allclose(res["proposals"][0], ores["proposals"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(res["samples"][0], ores["samples"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(res["logposts"][0], ores["logposts"], rtol=1e-8)
    np.testing.assert_allclose(res["accept_rates"][0], ores["accept_rates"], atol=1e-12)
    # ensemble 1 of the same launch == a separate launch with ensemble_offset=1
    res1 = s.pt_run(kind, p, q, nsamples, burnin, thin=thin, ntemps=ntemps, n_ensembles=1, seed=1234,
                    ensemble_offset=1, prior=pr)
    assert np.array_equal(res1["samples"][0], res["samples"][1])
    s.close()


def test_pt_run_with_init_and_parallel_mode(C, O):
    from carma_pack_b200 import synth
    t, y, e = synth.readme_series(90, 7)
    s = C.Series(t, y, e)
    pr = s.default_prior()
    init = synth.readme_theta(3)
    res = s.pt_run(C.KIND_CARMA, 5, 3, 40, 40, ntemps=5, n_ensembles=2, seed=9, init=init, prior=pr, record_trace=True)
    ores = O.pt_run(O.KIND_CARMA, 5, 3, t, y, e, 40, 40, ntemps=5, seed=9, init=init, prior=O.default_prior(t, y),
                    want_trace=True)
    # first proposals start from init for every chain
    assert np.array_equal(res["ram_trace"][0]["accepted"], ores["ram_trace"]["accepted"])
    np.testing.assert_allclose(res["samples"][0], ores["samples"], rtol=1e-7, atol=1e-9)
    # order_mode 1 (concurrent proposals) is a valid sampler: finite, and consistent stored log-posteriors
    r1 = s.pt_run(C.KIND_CARMA, 5, 3, 50, 50, ntemps=5, n_ensembles=4, seed=9, prior=pr, order_mode=1)
    relp = s.loglik(C.KIND_CARMA, 5, 3, r1["samples"].reshape(-1, 11), prior=pr).reshape(r1["logposts"].shape)
    np.testing.assert_allclose(relp, r1["logposts"], rtol=1e-8)
    s.close()


def test_multi_series_pt_run_equals_per_curve_runs(C):
    """Survey-scale MCMC: one launch over a ragged batch of light curves gives, for every curve, exactly the
    chain of a single-series run with the matching global ensemble index."""
    from carma_pack_b200 import synth
    rng = np.random.default_rng(31)
    ar, ma, s2 = synth.carma31_truth()
    ts, ys, es, off = [], [], [], [0]
    for c in range(5):
        ny = int(rng.integers(40, 160))
        t = np.cumsum(rng.uniform(0.5, 2.0, ny))
        y = 3.0 + synth.carma_process(t, s2, ar, ma, rng) + 0.1 * rng.standard_normal(ny)
        ts.append(t); ys.append(y); es.append(np.full(ny, 0.1)); off.append(off[-1] + ny)
    m = C.MultiSeries(np.concatenate(ts), np.concatenate(ys), np.concatenate(es), off)
    nens, T = 3, 4
    res = m.pt_run(C.KIND_CARMA, 3, 1, 30, 40, ntemps=T, n_ensembles=nens, seed=77)
    assert res["samples"].shape == (5, nens, 30, 7) and np.all(np.isfinite(res["logposts"]))
    for c in (0, 2, 4):
        s = C.Series(ts[c], ys[c], es[c])
        one = s.pt_run(C.KIND_CARMA, 3, 1, 30, 40, ntemps=T, n_ensembles=nens, seed=77, ensemble_offset=c * nens,
                       prior=s.default_prior())
        assert np.array_equal(one["samples"], res["samples"][c])
        assert np.array_equal(one["logposts"], res["logposts"][c])
        assert np.array_equal(one["accept_rates"], res["accept_rates"][c])
        s.close()
    m.close()


def test_pt_posterior_recovers_truth_car1(C):
    """Statistical check in the spirit of carma_unit_tests.cpp:1319-1375 (posterior within 3 sigma of truth),
    using many independent ensembles instead of one long chain."""
    from carma_pack_b200 import synth
    rng = np.random.default_rng(3)
    tau, sig_y, mu = 25.0, 1.5, 3.0
    t = np.cumsum(rng.uniform(0.5, 1.5, 300))
    y0 = mu + synth.car1_process(t, 2 * sig_y ** 2 / tau, tau, rng)
    e = np.full(t.size, 0.15)
    y = y0 + e * rng.standard_normal(t.size)
    s = C.Series(t, y, e)
    res = s.pt_run(C.KIND_CAR1, 1, 0, 400, 1500, thin=5, ntemps=1, n_ensembles=64, seed=5)
    smp = res["samples"].reshape(-1, 4)
    assert 0.1 < res["accept_rates"].mean() < 0.5  # RAM targets 0.25
    logw = smp[:, 3]
    assert abs(logw.mean() - np.log(1 / tau)) < 3 * logw.std()
    assert abs(smp[:, 2].mean() - mu) < 3 * smp[:, 2].std()
    assert abs(smp[:, 0].mean() - sig_y) < 3 * smp[:, 0].std()
    assert abs(smp[:, 1].mean() - 1.0) < 3 * smp[:, 1].std()
    # between-ensemble agreement (independent chains sample the same posterior)
    ens_means = res["samples"][:, :, 3].mean(axis=1)
    assert ens_means.std() < 3 * logw.std()
    s.close()


def test_simulated_survey_is_a_draw_from_the_model(C, O):
    """carma_multi_series_simulate: light curves generated in HBM.  (1) reproducible and independent of how the
    survey is sharded; (2) per-curve default priors equal the host recipe on the downloaded data; (3) K4 on the
    resident curves equals the oracle on the downloaded ones; (4) the oracle's standardized one-step residuals
    of the curves at theta_true are N(0,1): the generator draws from exactly the model the filter evaluates."""
    from carma_pack_b200 import synth
    th = synth.carma31_theta(sigmay=1.7, mu=4.0)
    ar, ma, s2 = synth.carma31_truth()
    nc, ny, yerr = 384, 400, 0.2
    m = C.MultiSeries.simulate(nc, ny, C.KIND_CARMA, 3, 1, th, yerr=yerr, dt_min=0.05, dt_max=60.0, seed=11)
    m2 = C.MultiSeries.simulate(7, ny, C.KIND_CARMA, 3, 1, th, yerr=yerr, dt_min=0.05, dt_max=60.0, seed=11, curve_offset=100)
    for c in range(7):
        a, b = m.curve(100 + c), m2.curve(c)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    m2.close()
    pri = m.default_priors()
    ts, ys, es, z_all = [], [], [], []
    for c in range(nc):
        t, y, e = m.curve(c)
        dt = np.diff(t)
        assert dt.min() >= 0.05 * (1 - 1e-9) and dt.max() <= 60.0 * (1 + 1e-9) and np.allclose(e, yerr, rtol=1e-15) and np.all(np.isfinite(y))
        ts.append(t); ys.append(y); es.append(e)
        op = O.default_prior(t, y)
        assert np.allclose(tuple(pri[c]), (op.max_stdev, op.max_freq, op.min_freq, op.kappa_low, op.kappa_high, 50.0),
                           rtol=1e-9), c
        if c < 96:
            mean, var = O.filterp(t, y - th[2], e, s2 * th[0] ** 2, ar, ma[:2])
            z_all.append((y - th[2] - mean) / np.sqrt(var))
    # Cauchy gaps: the median of dt_min + |C| is dt_min + 1
    assert abs(np.median(np.concatenate([np.diff(t) for t in ts])) - 1.05) < 0.03
    tcat, ycat, ecat = np.concatenate(ts), np.concatenate(ys), np.concatenate(es)
    off = np.arange(nc + 1) * ny
    thetas = np.tile(th, (nc, 1))
    thetas[:, 3:] += 0.1 * np.random.default_rng(2).standard_normal((nc, 4))
    got = m.loglik(C.KIND_CARMA, 3, 1, thetas)
    opri = [O.default_prior(ts[c], ys[c]) for c in range(nc)]
    want = O.logdensity_multi(O.KIND_CARMA, 3, 1, tcat, ycat, ecat, off, thetas, opri)
    assert_logpost_parity(got, want, what="simulated survey")
    z = np.concatenate(z_all)                       # 96 x 400 = 38,400 residuals
    assert abs(z.mean()) < 4.0 / np.sqrt(z.size)
    assert abs(z.var() - 1.0) < 4.0 * np.sqrt(2.0 / z.size)
    assert abs(np.mean(z[1:] * z[:-1])) < 4.0 / np.sqrt(z.size)   # white
    assert abs(np.mean(z ** 4) - 3.0) < 0.2                       # Gaussian tails
    # the variance of the process is sigma_y^2 (+ noise) at theta_true
    assert abs(np.var(ycat) / (1.7 ** 2 + yerr ** 2) - 1.0) < 0.1
    m.close()
    with pytest.raises(C.CarmaError):
        C.MultiSeries.simulate(4, 1, C.KIND_CARMA, 3, 1, th)


def test_simulated_survey_other_model_kinds(C, O):
    """The generator for CAR(1) and for a real-root CAR(4) (generic transition blocks): the K4 log-likelihood of the
    simulated curves at theta_true equals the oracle's, and its standardized residuals are N(0,1)."""
    # CAR(1): theta = (sigma_y, measerr_scale, mu, log omega)
    sig_y, mu, omega, yerr = 1.3, -2.0, 0.08, 0.1
    th = np.array([sig_y, 1.0, mu, np.log(omega)])
    m = C.MultiSeries.simulate(128, 300, C.KIND_CAR1, 1, 0, th, yerr=yerr, dt_min=0.2, dt_max=40.0, seed=5)
    zs = []
    for c in range(128):
        t, y, e = m.curve(c)
        mean, var = O.filter1(t, y - mu, e, 2.0 * sig_y ** 2 * omega, omega)
        zs.append((y - mu - mean) / np.sqrt(var))
    z = np.concatenate(zs)
    assert abs(z.mean()) < 4.0 / np.sqrt(z.size) and abs(z.var() - 1.0) < 4.0 * np.sqrt(2.0 / z.size)
    m.close()
    # CAR(4) with two overdamped (real-root) quadratic factors
    th4 = np.array([0.9, 1.0, 1.0, np.log(0.02), np.log(0.9), np.log(0.5), np.log(3.0)])
    m = C.MultiSeries.simulate(96, 250, C.KIND_CARP, 4, 0, th4, yerr=0.05, dt_min=0.1, dt_max=30.0, seed=6)
    ts, ys, es = zip(*[m.curve(c) for c in range(96)])
    off = np.arange(97) * 250
    thetas = np.tile(th4, (96, 1))
    got = m.loglik(C.KIND_CARP, 4, 0, thetas, flags=C.IGNORE_BOUNDS)
    opri = [O.default_prior(ts[c], ys[c]) for c in range(96)]
    want = O.logdensity_multi(O.KIND_CARP, 4, 0, np.concatenate(ts), np.concatenate(ys), np.concatenate(es), off, thetas,
                              opri, ignore_prior=True)
    assert_logpost_parity(got, want, rtol=1e-8, what="simulated CAR(4)")
    # -2 loglik + sum log var = chi^2 with ny dof per curve: the mean over curves of the quadratic form is ny
    roots = O.ar_roots(th4[3:7])
    sig2 = th4[0] ** 2 / O.variance(roots, np.array([1.0, 0.0, 0.0, 0.0]))
    chi2 = []
    for c in range(96):
        mean, var = O.filterp(ts[c], ys[c] - th4[2], es[c], sig2, roots, [1.0])
        chi2.append(np.sum((ys[c] - th4[2] - mean) ** 2 / var))
    chi2 = np.array(chi2)
    assert abs(chi2.mean() / 250.0 - 1.0) < 4.0 * np.sqrt(2.0 / (250.0 * 96.0))
    m.close()


def test_pt_helper_warp_kernel_is_bit_identical_to_the_plain_kernel(C):
    """Small launches run the warp-specialised PT kernel (producer warps compute the transition blocks, the chain warps
    only the state recursion; hand-over through a shared-memory ring).  Same operations on the same values: samples,
    log-posteriors, acceptance and exchange rates are bit-identical to the plain kernel, for one and for several
    ensembles, odd and even orders, CAR(1) (no exchanges) and a series with a real root pair among the chains."""
    import os
    from carma_pack_b200 import synth
    cases = [(C.KIND_CARMA, 5, 3, 270, 10, 1), (C.KIND_CARMA, 5, 3, 270, 10, 7), (C.KIND_CARP, 4, 0, 151, 6, 13),
             (C.KIND_CARMA, 7, 2, 90, 12, 3), (C.KIND_CAR1, 1, 0, 64, 1, 5), (C.KIND_ZCARMA, 3, 0, 120, 4, 20),
             (C.KIND_CARMA, 2, 1, 12, 3, 2)]
    for kind, p, q, ny, ntemps, nens in cases:
        t, y, e = synth.readme_series(max(ny, 2), 7 + ny)
        s = C.Series(t, y, e)
        out = {}
        for mode in ("0", "1"):
            os.environ["CARMA_PT_HELP"] = mode
            out[mode] = s.pt_run(kind, p, q, nsamples=40, burnin=60, ntemps=ntemps, n_ensembles=nens, seed=77 + p)
        os.environ.pop("CARMA_PT_HELP")
        for k in ("samples", "logposts", "accept_rates", "exchange_rates"):
            assert np.array_equal(out["0"][k], out["1"][k], equal_nan=True), (kind, p, q, ny, k)
        assert np.all(np.isfinite(out["1"]["logposts"]))
        s.close()


def test_pt_time_sliced_launch_is_bit_identical_to_one_block_per_group(C):
    """When the number of 64-chain groups is not what the GPU holds evenly, pt_kernel runs as W < groups worker blocks
    that pull (group, slice of ticks) units from a queue and park the chains in HBM between slices.  The chains do
    the same operations on the same values whichever worker runs a slice: samples, log-posteriors, acceptance and
    exchange rates and the recorded accept/exchange traces are bit-identical to the one-block-per-group launch --
    for few workers (every group migrates between SMs), one worker (fully sequential), more groups than workers
    by one, adaptation running across slice boundaries, and both step orders."""
    import os
    from carma_pack_b200 import synth
    cases = [(C.KIND_CARMA, 5, 3, 120, 10, 100, 0, (1, 7, 16)), (C.KIND_CARMA, 3, 1, 60, 4, 333, 1, (20,)),
             (C.KIND_CAR1, 1, 0, 64, 1, 200, 0, (3,)), (C.KIND_ZCARMA, 4, 0, 80, 6, 45, 0, (4,))]
    for kind, p, q, ny, ntemps, nens, order_mode, workers in cases:
        t, y, e = synth.readme_series(ny, 11 + ny)
        s = C.Series(t, y, e)
        os.environ["CARMA_PT_SLICE"] = "0"
        os.environ["CARMA_PT_HELP"] = "0"
        ref = s.pt_run(kind, p, q, nsamples=30, burnin=50, ntemps=ntemps, n_ensembles=nens, seed=5 + p, order_mode=order_mode)
        for w in workers:
            os.environ["CARMA_PT_SLICE"] = str(w)
            got = s.pt_run(kind, p, q, nsamples=30, burnin=50, ntemps=ntemps, n_ensembles=nens, seed=5 + p, order_mode=order_mode)
            for k in ("samples", "logposts", "accept_rates", "exchange_rates"):
                assert np.array_equal(ref[k], got[k], equal_nan=True), (kind, p, q, w, k)
        os.environ.pop("CARMA_PT_SLICE")
        os.environ.pop("CARMA_PT_HELP")
        assert np.all(np.isfinite(ref["logposts"]))
        s.close()
    # the recorded traces too (small case), and the automatic choice on a launch larger than one wave of blocks
    t, y, e = synth.readme_series(40, 3)
    s = C.Series(t, y, e)
    os.environ["CARMA_PT_HELP"] = "0"
    os.environ["CARMA_PT_SLICE"] = "0"
    a = s.pt_run(C.KIND_CARMA, 2, 1, nsamples=10, burnin=20, ntemps=3, n_ensembles=50, seed=9, record_trace=True)
    os.environ["CARMA_PT_SLICE"] = "2"
    b = s.pt_run(C.KIND_CARMA, 2, 1, nsamples=10, burnin=20, ntemps=3, n_ensembles=50, seed=9, record_trace=True)
    for k in a:
        if isinstance(a[k], np.ndarray):
            if a[k].dtype.names:
                for f in a[k].dtype.names:
                    if f != "pad":
                        assert np.array_equal(a[k][f], b[k][f], equal_nan=True), (k, f)
            else:
                assert np.array_equal(a[k], b[k], equal_nan=True), k
    os.environ["CARMA_PT_SLICE"] = "0"
    c0 = s.pt_run(C.KIND_CARMA, 2, 1, nsamples=5, burnin=20, ntemps=3, n_ensembles=21 * 1000, seed=9)
    os.environ.pop("CARMA_PT_SLICE")
    c1 = s.pt_run(C.KIND_CARMA, 2, 1, nsamples=5, burnin=20, ntemps=3, n_ensembles=21 * 1000, seed=9)   # 1,000 groups: automatic
    os.environ.pop("CARMA_PT_HELP")
    assert np.array_equal(c0["samples"], c1["samples"]) and np.array_equal(c0["exchange_rates"], c1["exchange_rates"])
    s.close()
