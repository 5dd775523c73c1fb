"""API-surface conformance on the GPU, mirroring the reference's src/tests/testCarmcmc.py
(10-point series, every overload arity, stored log-posterior == getLogDensity(sample), Predict
variance grows when extrapolating) plus the CarmaModel drivers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cm():
    import carma_pack_b200 as c
    if c._lib.device_count() < 1:
        pytest.fail("no CUDA device visible")
    from carma_pack_b200 import _carmcmc
    c._carmcmc_mod = _carmcmc
    return c


@pytest.fixture(scope="module")
def small(cm):
    # testCarmcmc.py:13-34
    rng = np.random.default_rng(1)
    npts = 10
    x = 1.0 * np.arange(npts)
    ar_roots = np.array([-0.06283185 - 1.25663706j, -0.06283185 + 1.25663706j, -0.02094395 - 0.25132741j,
                         -0.02094395 + 0.25132741j, -0.03141593 + 0.j])
    sigsqr = 0.00126811439419
    y = cm.carma_process(x, sigsqr, ar_roots, rng=rng)
    dy = np.sqrt(sigsqr) * np.ones(npts)
    m = cm._carmcmc_mod
    xd, yd, dyd = m.vecD(), m.vecD(), m.vecD()
    xd.extend(x); yd.extend(y); dyd.extend(dy)
    return dict(x=x, y=y, dy=dy, xd=xd, yd=yd, dyd=dyd, nSample=100, nBurnin=10, nThin=1, nWalkers=2)


def test_car1_all_arities(cm, small):  # testCarmcmc.py:36-52
    m = cm._carmcmc_mod
    s = small
    m.set_seed(11)
    cpp = m.run_mcmc_car1(s["nSample"], s["nBurnin"], s["xd"], s["yd"], s["dyd"], s["nThin"])
    psamples = np.array(cpp.getSamples())
    plog = np.array(cpp.GetLogLikes())
    assert psamples.shape == (100, 4) and plog.shape == (100,)
    sample0 = m.vecD(); sample0.extend(psamples[0])
    assert np.isfinite(cpp.getLogPrior(sample0))
    assert abs(plog[0] - cpp.getLogDensity(sample0)) < 1e-7 * max(1, abs(plog[0]))
    m.run_mcmc_car1(s["nSample"], s["nBurnin"], s["xd"], s["yd"], s["dyd"])
    guess = cpp.getSamples()[0]
    m.run_mcmc_car1(s["nSample"], s["nBurnin"], s["xd"], s["yd"], s["dyd"], s["nThin"], guess)


@pytest.mark.parametrize("p,q", [(3, 0), (3, 2), (3, 1)])
def test_carma_all_arities(cm, small, p, q):  # testCarmcmc.py:54-104
    m = cm._carmcmc_mod
    s = small
    m.set_seed(5)
    a = (s["nSample"], s["nBurnin"], s["xd"], s["yd"], s["dyd"], p, q, s["nWalkers"])
    m.run_mcmc_carma(*a)
    m.run_mcmc_carma(*a, False)
    sampler = m.run_mcmc_carma(*a, False, s["nThin"])
    psamples = np.array(sampler.getSamples())
    plog = np.array(sampler.GetLogLikes())
    assert psamples.shape == (100, 3 + p + q)
    sample0 = m.vecD(); sample0.extend(psamples[0])
    assert abs(plog[0] - sampler.getLogDensity(sample0)) < 1e-7 * max(1, abs(plog[0]))
    guess = sampler.getSamples()[0]
    again = m.run_mcmc_carma(*a, False, s["nThin"], guess)
    assert np.array(again.getSamples()).shape == psamples.shape
    # SetMLE(True): bounds ignored, prior term still added
    sampler.SetMLE(True)
    assert np.isfinite(sampler.getLogDensity(sample0))
    # numpy arrays / lists are accepted where the reference needs vecD
    assert abs(sampler.getLogDensity(list(psamples[0])) - sampler.getLogDensity(sample0)) == 0.0
    lp = np.array(sampler.getLogDensityBatch(m.vecvecD([m.vecD(list(r)) for r in psamples[:5]])))
    assert np.allclose(lp, [sampler.getLogDensity(list(r)) for r in psamples[:5]], rtol=1e-12)
    # wrong-length init falls back to prior draws with a warning (carpack.cpp:481-484)
    m.run_mcmc_carma(*a, False, 1, m.vecD([1.0, 2.0]))


def test_zcarma_flag_runs_zcar(cm, small):
    m = cm._carmcmc_mod
    s = small
    sampler = m.run_mcmc_carma(50, 10, s["xd"], s["yd"], s["dyd"], 3, 0, 2, True)
    assert np.array(sampler.getSamples()).shape == (50, 6)


def test_kalman1_predict(cm, small):  # testCarmcmc.py:106-117
    m = cm._carmcmc_mod
    s = small
    kf = m.KalmanFilter1(s["xd"], s["yd"], s["dyd"], 1.0, 1.0)
    kf.Filter()
    assert len(kf.GetMean()) == 10 and len(kf.GetVar()) == 10
    pred0 = kf.Predict(s["xd"][0])
    predN = kf.Predict(s["xd"][-1] + 1)
    assert predN.second > pred0.second
    sim = np.array(kf.Simulate(m.vecD([0.5, 3.3, 12.0])))
    assert sim.shape == (3,) and np.all(np.isfinite(sim))


def test_kalmanp_predict_and_golden(cm, small, kelly):  # testCarmcmc.py:119-146
    m = cm._carmcmc_mod
    s = small
    sampler = m.run_mcmc_carma(s["nSample"], s["nBurnin"], s["xd"], s["yd"], s["dyd"], 4, 0, s["nWalkers"], False, 1)
    trace = np.array(sampler.getSamples())
    smp = cm.CarmaSample(s["x"], s["y"], s["dy"], trace=trace, logpost=np.array(sampler.GetLogLikes()), p=4, q=0)
    sigsqr = float(smp._samples["sigma"][0][0]) ** 2
    omega = m.vecC()
    for r in smp._samples["ar_roots"][0]:
        omega.append(complex(r))
    ma = m.vecD([1.0, 0.0, 0.0, 0.0])
    kf = m.KalmanFilterp(s["xd"], s["yd"], s["dyd"], sigsqr, omega, ma)
    kf.Filter()
    pred0 = kf.Predict(s["xd"][0])
    predN = kf.Predict(s["xd"][-1] + 1)
    assert predN.second > pred0.second
    # the class API reproduces the golden filter of the reference (carma_unit_tests.cpp:387-444)
    kf = m.KalmanFilterp(m.vecD(kelly["t"]), m.vecD(kelly["y"]), m.vecD(kelly["yerr"]), float(kelly["sigsqr"]),
                         m.vecC([complex(z) for z in kelly["roots"]]), m.vecD(kelly["ma"]))
    kf.Filter()
    np.testing.assert_allclose(np.array(kf.GetVar()), kelly["var"], rtol=1e-9)
    np.testing.assert_allclose(np.array(kf.GetMean()), kelly["mean"], rtol=0, atol=1e-9)
    pm, pv = kf.PredictMany(m.vecD(kelly["predict_t"]))
    np.testing.assert_allclose(np.array(pm), kelly["predict_mean"], rtol=1e-7, atol=1e-9)
    # unsorted input with a duplicate time is sorted and de-duplicated like KalmanFilter::init (kfilter.hpp:43-76)
    t2 = np.concatenate([kelly["t"][:50][::-1], kelly["t"][:1]])
    y2 = np.concatenate([kelly["y"][:50][::-1], kelly["y"][:1]])
    e2 = np.concatenate([kelly["yerr"][:50][::-1], kelly["yerr"][:1]])
    kf2 = m.KalmanFilterp(m.vecD(t2), m.vecD(y2), m.vecD(e2), float(kelly["sigsqr"]),
                          m.vecC([complex(z) for z in kelly["roots"]]), m.vecD(kelly["ma"]))
    kf2.Filter()
    assert len(kf2.GetVar()) == 50
    np.testing.assert_allclose(np.array(kf2.GetVar()), kelly["var"][:50], rtol=1e-9)


def test_carma_model_run_mcmc_and_sample(cm):
    from carma_pack_b200 import synth
    t, y, e = synth.readme_series(120, 21)
    model = cm.CarmaModel(t, y, e, p=3, q=1)
    sample = model.run_mcmc(300, nburnin=300, seed=3, n_ensembles=4)
    assert sample._samples["logpost"].shape == (1200, 1)
    for k in ("ar_roots", "psd_centroid", "psd_width", "ar_coefs", "ma_coefs", "sigma", "var", "mu", "loglik"):
        assert k in sample.parameters
    # loglik (SetMLE) = logpost for in-prior samples: same prior term is added.  CarmaSample takes it from
    # the stored log-posteriors; the reference's re-filtering loop (one batched launch here) gives the same numbers
    assert np.array_equal(sample._samples["loglik"], sample._samples["logpost"])
    np.testing.assert_allclose(sample.recompute_loglik(), np.ravel(sample._samples["logpost"]), rtol=1e-8)
    tq = np.linspace(t[0] - 5, t[-1] + 20, 64)
    pm, pv = sample.predict(tq)
    assert pm.shape == (64,) and np.all(pv > 0) and pv[-1] > pv[len(pv) // 2]
    mean, var = sample.kalman_filter()
    chi = (y - mean) / np.sqrt(var)
    assert 0.5 < chi.std() < 2.0  # standardized residuals are O(1) (assess_fit, carma_pack.py:687-744)
    assert np.isfinite(sample.DIC())
    resid, acf1, acf2, bound = sample.assess_fit_values()
    assert resid.shape == y.shape and abs(acf1[0] - 1.0) < 1e-12 and np.mean(np.abs(acf1[1:]) < 3 * bound) > 0.8
    lo, hi, mid, freq = sample.psd_credible_band(percentile=95.0, nsamples=200)
    assert np.all(lo <= mid) and np.all(mid <= hi) and freq.shape == mid.shape
    # power_spectrum() of the MAP sample sits inside its own posterior band at most frequencies
    i0 = sample.best_index()
    psd_map = cm.power_spectrum(freq, float(sample._samples["sigma"][i0][0]), sample._samples["ar_coefs"][i0],
                                sample._samples["ma_coefs"][i0])
    assert np.mean((psd_map >= lo) & (psd_map <= hi)) > 0.7
    ysim = sample.simulate(np.array([t[-1] + 1.0, t[10] + 0.01, t[-1] + 5.0]), seed=4)
    assert ysim.shape == (3,) and np.all(np.isfinite(ysim))
    # CAR(1) path
    m1 = cm.CarmaModel(t, y, e, p=1)
    s1 = m1.run_mcmc(200, seed=4)
    assert s1._samples["log_omega"].shape == (200, 1)


def test_choose_order_small_grid(cm):
    """choose_order over p <= 2: AICc bookkeeping as carma_pack.py:173-190 and MLEs at least as good as
    the best random start; the GPU MLE log-likelihood is reproduced by the CPU oracle."""
    from carma_pack_b200 import synth
    from oracle import oracle as O
    rng = np.random.default_rng(2)
    t = np.cumsum(rng.uniform(0.5, 1.5, 200))
    y = 1.0 + synth.car1_process(t, 2 * 1.2 ** 2 / 8.0, 8.0, rng)
    e = np.full(t.size, 0.1)
    y = y + e * rng.standard_normal(t.size)
    model = cm.CarmaModel(t, y, e)
    mle, pqlist, aicc = model.choose_order(2, ntrials=24, seed=7, verbose=False)
    assert pqlist == [(1, 0), (2, 0), (2, 1)]
    assert (model.p, model.q) == pqlist[int(np.argmin(aicc))]
    assert all(np.isfinite(aicc))
    # CAR(1) data: CAR(1) is competitive and its MLE sits near the truth
    m1 = model.get_mle(1, 0, ntrials=24, seed=8)
    assert abs(m1.x[3] - np.log(1 / 8.0)) < 1.0
    pr = O.default_prior(t, y)
    want = O.logdensity(O.KIND_CAR1, 1, 0, t, y, e, m1.x, prior=pr)[0]
    assert abs(-m1.fun - want) < 1e-8 * abs(want)
    k = 2 + 1
    assert abs(aicc[0] - (2 * k + 2 * model.get_mle(1, 0, ntrials=24, seed=7).fun + 2 * k * (k + 1) / (t.size - k - 1))) < 0.5


def test_reference_package_name_and_flows(cm, small):
    """`import carmcmc` keeps working, and the bodies of the reference's testCarp / testCarpq / testCar1
    (src/tests/testCarmcmc.py:36-104) run unchanged against it."""
    import carmcmc
    s = small
    xdata, ydata, dydata = carmcmc.vecD(), carmcmc.vecD(), carmcmc.vecD()
    xdata.extend(s["x"]); ydata.extend(s["y"]); dydata.extend(s["dy"])
    # testCar1
    cppSample = carmcmc.run_mcmc_car1(100, 10, xdata, ydata, dydata, 1)
    psampler = carmcmc.Car1Sample(s["x"], s["y"], s["dy"], cppSample)
    assert psampler.p == 1
    psamples = np.array(cppSample.getSamples())
    ploglikes = np.array(cppSample.GetLogLikes())
    sample0 = carmcmc.vecD(); sample0.extend(psamples[0])
    assert np.isfinite(cppSample.getLogPrior(sample0))
    assert round(abs(ploglikes[0] - cppSample.getLogDensity(sample0)), 7) == 0
    # testCarp (pModel=3) and testCarpq (pModel=3, qModel=2): note the reference's own p = p+q quirk
    for pModel, qModel in [(3, 0), (3, 2)]:
        sampler = carmcmc.run_mcmc_carma(100, 10, xdata, ydata, dydata, pModel, qModel, 2, False, 1)
        psampler = carmcmc.CarmaSample(np.array(xdata), np.array(ydata), np.array(dydata), sampler)
        assert psampler.p == pModel + qModel
        psamples = np.array(sampler.getSamples())
        ploglikes = np.array(sampler.GetLogLikes())
        sample0 = carmcmc.vecD(); sample0.extend(psamples[0])
        assert round(abs(ploglikes[0] - sampler.getLogDensity(sample0)), 7) == 0
    # README flow
    model = carmcmc.CarmaModel(s["x"], s["y"], s["dy"], p=2, q=0)
    smp = model.run_mcmc(50)
    assert smp.get_samples("sigma").shape == (50, 1)
    assert carmcmc.get_ar_roots(np.array([0.01]), np.array([0.2])).shape == (2,)


def test_simulate_is_conditional_draw(cm, kelly):
    """KalmanFilter::Simulate (kfilter.hpp:135-184): draws at one new time follow N(Predict mean, Predict var);
    two simulated times are positively correlated when close (the first draw is inserted before the second)."""
    m = cm._carmcmc_mod
    n = 80
    t, y, e = kelly["t"][:n], kelly["y"][:n], kelly["yerr"][:n]
    kf = m.KalmanFilterp(m.vecD(t), m.vecD(y), m.vecD(e), float(kelly["sigsqr"]),
                         m.vecC([complex(z) for z in kelly["roots"]]), m.vecD(kelly["ma"]))
    tq = 0.5 * (t[30] + t[31])
    pred = kf.Predict(tq)
    m.set_seed(99)
    draws = np.array([kf.Simulate(m.vecD([tq]))[0] for _ in range(300)])
    assert abs(draws.mean() - pred.first) < 4 * np.sqrt(pred.second / 300)
    assert 0.7 < draws.var() / pred.second < 1.4
    pairs = np.array([list(kf.Simulate(m.vecD([tq, tq + 0.2]))) for _ in range(200)])
    assert np.corrcoef(pairs[:, 0], pairs[:, 1])[0, 1] > 0.5
    # the filter object is restored afterwards
    kf.Filter()
    np.testing.assert_allclose(np.array(kf.GetVar()), kelly["var"][:n], rtol=1e-9)


def test_native_optimizer_matches_numpy_twin(cm):
    """carma_mle_batch (C++ host loop) against batched_lbfgs (numpy) from the same starts: same algorithm, so the
    best optimum agrees closely and the per-start optima agree for the large majority of starts (summation order
    differs, and forward-difference gradients amplify that on flat directions)."""
    from carma_pack_b200 import synth
    t, y, e = synth.readme_series(200, 11)
    model = cm.CarmaModel(t, y, e)
    for p, q in ((1, 0), (3, 1), (5, 2)):
        a = model.get_mle(p, q, ntrials=32, seed=21, optimizer="native")
        b = model.get_mle(p, q, ntrials=32, seed=21, optimizer="python")
        assert np.isfinite(a.fun) and np.isfinite(b.fun)
        assert abs(a.fun - b.fun) < 1e-3 * max(1.0, abs(b.fun)), (p, q, a.fun, b.fun)
        fa, fb = np.asarray(a.all_fun), np.asarray(b.all_fun)
        both = (fa < 1e299) & (fb < 1e299)
        assert both.sum() >= 24
        close = np.abs(fa[both] - fb[both]) < 1e-2 * np.maximum(1.0, np.abs(fb[both]))
        assert close.mean() > 0.7, (p, q, close.mean())
        # every native optimum is a genuine function value: re-evaluate it
        kind = cm.KIND_CAR1 if p == 1 else (cm.KIND_CARMA if q > 0 else cm.KIND_CARP)
        flags = 0 if p == 1 else cm.IGNORE_BOUNDS
        re = -model.series.loglik(kind, p, q, a.all_x[both], prior=model.series.default_prior(True), flags=flags)
        np.testing.assert_allclose(re, fa[both], rtol=1e-12)
        assert a.nfev > 0 and a.nit >= 1
    # argument checking
    with pytest.raises(cm.CarmaError):
        model.series.mle_batch(cm.KIND_CARMA, 3, 1, np.zeros((2, 7)), -np.ones(7), np.ones(7), slot=5)


def test_device_optimizer_matches_the_host_loop(cm):
    """carma_mle_batch_device (the whole fit of a start inside one kernel, a warp per start) against carma_mle_batch
    (host loop) from the same starts.  Same algorithm decision for decision, the trial points evaluated by the same
    arithmetic as K1 (shared-memory LU prologue) and lane 0's optimiser arithmetic rounded like the host compiler's:
    the two take the same path, so every start ends at the same point with the same value after the same number of
    iterations -- bit for bit.  (The evaluation counts differ: the device tries every halving of a step that fits
    into the warp's free lanes at once, the host loop four per launch.)"""
    from carma_pack_b200 import synth
    t, y, e = synth.readme_series(200, 11)
    model = cm.CarmaModel(t, y, e)
    for p, q in ((1, 0), (2, 0), (3, 1), (5, 2), (7, 4)):
        a = model.get_mle(p, q, ntrials=32, seed=21, optimizer="native")
        b = model.get_mle(p, q, ntrials=32, seed=21, optimizer="device")
        assert np.isfinite(a.fun) and np.isfinite(b.fun)
        fa, fb = np.asarray(a.all_fun), np.asarray(b.all_fun)
        np.testing.assert_array_equal(fa, fb)
        np.testing.assert_array_equal(np.asarray(a.all_x), np.asarray(b.all_x))
        assert a.nit == b.nit and a.fun == b.fun
        both = fb < 1e299
        assert both.sum() >= 24
        # every optimum is a genuine function value at a point inside the box
        kind = cm.KIND_CAR1 if p == 1 else (cm.KIND_CARMA if q > 0 else cm.KIND_CARP)
        flags = 0 if p == 1 else cm.IGNORE_BOUNDS
        re = -model.series.loglik(kind, p, q, b.all_x[both], prior=model.series.default_prior(True), flags=flags)
        np.testing.assert_array_equal(re, fb[both])
        assert b.nfev > 0 and b.nit >= 1
    # a start's fit does not depend on the other starts of the launch
    kind, x0, lo, hi, prior, flags = model.mle_starts(3, 1, 16, seed=5)
    xa, fa, _, _ = model.series.mle_batch(kind, 3, 1, x0, lo, hi, prior=prior, flags=flags, on_device=True)
    xb, fb, _, _ = model.series.mle_batch(kind, 3, 1, x0[5:9], lo, hi, prior=prior, flags=flags, on_device=True)
    np.testing.assert_array_equal(fa[5:9], fb)
    np.testing.assert_array_equal(xa[5:9], xb)
    with pytest.raises(cm.CarmaError):
        model.series.mle_batch(kind, 3, 1, x0, lo, hi, prior=prior, flags=flags, history=9, on_device=True)
    # several models in one launch (carma_mle_grid_device): every job's result equals its own single-model launch
    jobs = [model.mle_starts(p, q, n, seed=40 + p) for p, q, n in ((1, 0, 5), (4, 2, 9), (2, 1, 12), (6, 0, 7), (3, 0, 3))]
    jobs[-1] = (jobs[-1][0], jobs[-1][1][:0]) + tuple(jobs[-1][2:])   # a job without starts
    pqs = ((1, 0), (4, 2), (2, 1), (6, 0), (3, 0))
    grid = model.series.mle_grid([(j[0], p, q) + tuple(j[1:]) for j, (p, q) in zip(jobs, pqs)])
    assert len(grid) == len(jobs)
    for (kind, x0, lo, hi, prior, flags), (p, q), (xg, fg, nitg, nfevg) in zip(jobs, pqs, grid):
        if x0.shape[0] == 0:
            assert xg.shape == (0, x0.shape[1]) and fg.size == 0 and nitg == 0 and nfevg == 0
            continue
        x1, f1, nit1, nfev1 = model.series.mle_batch(kind, p, q, x0, lo, hi, prior=prior, flags=flags, on_device=True)
        np.testing.assert_array_equal(xg, x1)
        np.testing.assert_array_equal(fg, f1)
        assert (nitg, nfevg) == (nit1, nfev1)
    assert model.series.mle_grid([]) == []


def test_device_optimizer_on_a_series_too_long_for_shared_memory(cm):
    """ny = 3000: the series plus the warps' work areas exceed the kernel's shared memory, the trial points are
    evaluated from global memory -- still the host loop's fits bit for bit."""
    from carma_pack_b200 import synth
    t, y, e = synth.readme_series(3000, 5)
    model = cm.CarmaModel(t, y, e)
    a = model.get_mle(3, 1, ntrials=8, seed=3, optimizer="native")
    b = model.get_mle(3, 1, ntrials=8, seed=3, optimizer="device")
    assert np.isfinite(b.fun) and a.nit == b.nit
    np.testing.assert_array_equal(np.asarray(a.all_fun), np.asarray(b.all_fun))
    np.testing.assert_array_equal(np.asarray(a.all_x), np.asarray(b.all_x))


def test_fast_filter_and_predict_equal_the_general_complex_kernels(cm):
    """KalmanFilterp::Filter / Predict for conjugate-symmetric roots run in the real-half recursion (time-parallel
    forward filter + per-query coefficient pass resuming from stored states); the general complex kernels
    (CARMA_PREDICT_GENERAL=1) are the same algorithm written as in kfilter.cpp:138-337.  Both must agree to rounding,
    for interpolation, forecasts, backcasts, query times that coincide with data times, and real-pair roots."""
    import os
    import time as _time
    from carma_pack_b200 import synth, Series
    rng = np.random.default_rng(12)
    for ny, p, q in ((400, 5, 3), (150, 4, 1), (90, 3, 0), (60, 2, 1), (40, 1, 0)):
        t, y, e = synth.readme_series(ny, 40 + ny)
        s = Series(t, y, e)
        if p == 1:
            roots, ma, sigsqr = np.array([-0.05 + 0j]), np.array([1.0]), 0.3
        else:
            th = synth.prior_draws(1, p, q, t, y, rng)[0]
            if p >= 4:
                th[3], th[4] = np.log(0.02), np.log(0.9)          # first factor: two real roots
            cs = cm.CarmaSample(t, y, e, trace=th[None, :], logpost=np.zeros(1), p=p, q=q)
            sigsqr, roots, ma, mu, scale = cs._params_at(0)
        tq = np.concatenate([np.linspace(t[0] - 0.1 * (t[-1] - t[0]), t[-1] + 0.2 * (t[-1] - t[0]), 257), t[::7], [t[0], t[-1]]])
        out = {}
        for mode in ("0", "1"):
            os.environ["CARMA_PREDICT_GENERAL"] = mode
            t0 = _time.perf_counter()
            out[mode] = (s.filter(sigsqr, roots, ma, measerr_scale=1.1, mu=0.3), s.predict(sigsqr, roots, ma, tq, measerr_scale=1.1, mu=0.3),
                         _time.perf_counter() - t0)
        os.environ.pop("CARMA_PREDICT_GENERAL")
        (m0, v0), (qm0, qv0), _ = out["0"]
        (m1, v1), (qm1, qv1), _ = out["1"]
        np.testing.assert_allclose(m0, m1, rtol=1e-9, atol=1e-9 * np.abs(y).max())
        np.testing.assert_allclose(v0, v1, rtol=1e-9)
        np.testing.assert_allclose(qm0, qm1, rtol=1e-7, atol=1e-8 * np.abs(y).max())
        np.testing.assert_allclose(qv0, qv1, rtol=1e-7)
        s.close()


def test_device_simulate_matches_the_exact_gaussian_process_conditional(cm):
    """carma_simulate (KalmanFilter<>::Simulate, kfilter.hpp:135-184; reference test carma_unit_tests.cpp:651-780):
    4,000 conditional paths in ONE call on a 60-point series.  The draws, whitened with the exact conditional mean and
    covariance of the Gaussian process (dense ny x ny algebra with carma_variance as the kernel), must be white
    N(0, 1): sample mean and covariance of the whitened draws within Monte-Carlo error, for times inside gaps, on top
    of data times, before the first and after the last point."""
    from carma_pack_b200 import synth, Series
    rng = np.random.default_rng(5)
    roots = synth.get_ar_roots(np.array([0.05, 0.02]), np.array([0.12]))      # CARMA(3,1): one pair + one real root
    ma = np.array([1.0, 2.5, 0.0])
    sigsqr = 1.7 ** 2 / synth.carma_variance(1.0, roots, ma)
    ny = 60
    t = np.cumsum(rng.uniform(0.5, 4.0, ny))
    y0 = synth.carma_process(t, sigsqr, roots, ma, rng=rng)
    e = rng.uniform(0.1, 0.4, ny)
    y = 3.0 + y0 + e * rng.standard_normal(ny)
    mu, scale = 3.0, 1.2
    tsim = np.concatenate([[t[0] - 7.0, t[0] - 1.0], t[5] + np.array([0.1, 0.3, 1.0]), [t[20]], rng.uniform(t[0], t[-1], 14), [t[-1] + 2.0, t[-1] + 30.0]])
    s = Series(t, y, e)
    npaths = 4000
    sims = s.simulate(sigsqr, roots, ma, tsim, measerr_scale=scale, mu=mu, seed=99, npaths=npaths)
    assert sims.shape == (npaths, tsim.size) and np.all(np.isfinite(sims))
    # reproducible, and path j does not depend on how many paths are drawn
    again = s.simulate(sigsqr, roots, ma, tsim, measerr_scale=scale, mu=mu, seed=99, npaths=7)
    assert np.array_equal(again, sims[:7])
    # exact conditional law
    kern = lambda a, b: np.array([[synth.carma_variance(sigsqr, roots, ma, lag=abs(x - z)) for z in b] for x in a])
    Kdd = kern(t, t) + np.diag(scale * e ** 2)
    Ksd = kern(tsim, t)
    Kss = kern(tsim, tsim)
    sol = np.linalg.solve(Kdd, (y - mu))
    cmean = mu + Ksd @ sol
    ccov = Kss - Ksd @ np.linalg.solve(Kdd, Ksd.T)
    # the conditional mean is also what Predict returns
    pm, pv = s.predict(sigsqr, roots, ma, tsim, measerr_scale=scale, mu=mu)
    np.testing.assert_allclose(pm + mu, cmean, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(pv, np.diag(ccov), rtol=1e-6, atol=1e-10)
    s.close()
    # the draw at a data time is NOT the datum (measurement noise): the conditional variance there is > 0
    w, V = np.linalg.eigh(ccov)
    keep = w > 1e-9 * w.max()
    white = (sims - cmean) @ V[:, keep] / np.sqrt(w[keep])
    n = npaths
    assert np.all(np.abs(white.mean(axis=0)) < 4.5 / np.sqrt(n)), np.abs(white.mean(axis=0)).max()
    C_ = np.cov(white.T)
    assert np.all(np.abs(np.diag(C_) - 1.0) < 5.0 * np.sqrt(2.0 / n)), np.abs(np.diag(C_) - 1.0).max()
    off = C_ - np.diag(np.diag(C_))
    assert np.abs(off).max() < 5.0 / np.sqrt(n), np.abs(off).max()


def test_device_post_processing_matches_reference_values_and_the_numpy_twin(cm, loglik_cases):
    """carma_derived_params (one thread per stored sample) against (i) the values the REFERENCE's own CarmaSample code
    gives for the same theta rows (tests/golden/derived_params.npz) and (ii) the vectorised numpy twin, for every golden
    (p, q) family incl. real-pair roots; ZCARMA rows (free kappa) through the C ABI directly against the golden values;
    and a 200,000-sample trace in one launch."""
    import os
    from conftest import GOLDEN, golden_case_names
    from carma_pack_b200 import _lib as L
    derived = dict(np.load(os.path.join(GOLDEN, "derived_params.npz")))
    t, y, e = loglik_cases["t60"], loglik_cases["y60"], loglik_cases["ysig60"]
    checked = zc = 0
    names = list(golden_case_names(loglik_cases))
    for name in names + [n for n in ("z5",) if n not in names]:
        p, q = int(loglik_cases[name + "_p"]), int(loglik_cases[name + "_q"])
        th = loglik_cases[name + "_theta"]
        if th.shape[1] != 3 + p + q:
            # ZCARMA (not a CarmaSample trace in the reference): theta carries kappa, the prior supplies its bounds;
            # expected values from the closed forms (carpack.cpp:687-698) and the reference's carma_variance
            from math import comb
            from carma_pack_b200 import synth
            s = L.Series(t, y, e)
            pr = s.default_prior()
            der = L.derived_params(L.KIND_ZCARMA, p, 0, th, prior=pr)
            s.close()
            tw = cm.CarmaSample(t, y, e, trace=th[:, :3 + p], logpost=np.zeros(th.shape[0]), p=p, q=0, postprocess="numpy")
            np.testing.assert_allclose(der["ar_roots"], tw._samples["ar_roots"], rtol=1e-13)
            kappa = (pr.kappa_high - pr.kappa_low) / (1.0 + np.exp(-th[:, 3 + p])) + pr.kappa_low
            ma = np.array([[comb(p - 1, i) / k ** i for i in range(p)] for k in kappa])
            np.testing.assert_allclose(der["ma_coefs"], ma, rtol=1e-12)
            for r in range(th.shape[0]):
                v1 = synth.carma_variance(1.0, der["ar_roots"][r], ma[r])
                np.testing.assert_allclose(der["sigma"][r] ** 2, th[r, 0] ** 2 / v1, rtol=1e-9)
            zc += 1
            continue
        want = derived[name + "_roots"]
        if True:
            lp = loglik_cases[name + "_logpost"]
            dev = cm.CarmaSample(t, y, e, trace=th, logpost=lp, p=p, q=q)               # default: device
            twin = cm.CarmaSample(t, y, e, trace=th, logpost=lp, p=p, q=q, postprocess="numpy")
            assert dev._series_obj is None                                              # no light curve is uploaded for this
            for k in ("ar_roots", "ar_coefs", "ma_coefs", "sigma", "psd_width", "psd_centroid"):
                a, b = np.asarray(dev._samples[k]), np.asarray(twin._samples[k])
                assert a.shape == b.shape or a.size == b.size, (name, k, a.shape, b.shape)
                np.testing.assert_allclose(np.ravel(a), np.ravel(b), rtol=1e-11, atol=1e-300, err_msg="%s %s" % (name, k))
            der = {"ar_roots": dev._samples["ar_roots"], "sigma": np.ravel(dev._samples["sigma"]),
                   "ma_coefs": np.pad(dev._samples["ma_coefs"], ((0, 0), (0, p - dev._samples["ma_coefs"].shape[1])))}
        np.testing.assert_allclose(der["ar_roots"], want, rtol=1e-13, atol=0, err_msg=name)
        np.testing.assert_allclose(der["ma_coefs"], derived[name + "_ma"], rtol=1e-12, atol=1e-15, err_msg=name)
        np.testing.assert_allclose(der["sigma"] ** 2, derived[name + "_sigsqr"], rtol=1e-10, err_msg=name)
        checked += 1
    assert checked >= 12 and zc >= 1
    rng = np.random.default_rng(1)
    big = np.tile(loglik_cases["c53_theta"][:1], (200000, 1)) + 1e-3 * rng.standard_normal((200000, 11))
    d = L.derived_params(L.KIND_CARMA, 5, 3, big)
    tw = cm.CarmaSample(t, y, e, trace=big[:500], logpost=np.zeros(500), p=5, q=3, postprocess="numpy")
    np.testing.assert_allclose(d["sigma"][:500], np.ravel(tw._samples["sigma"]), rtol=1e-10)
    assert np.all(np.isfinite(d["sigma"])) and d["ar_roots"].shape == (200000, 5)
