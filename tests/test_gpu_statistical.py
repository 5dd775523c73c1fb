"""The reference's long statistical tests, on its own fixtures, through the GPU path; and the choose_order parity
tests of BASELINE config 4.

  * residual whiteness (Anderson-Darling + ACF of residuals and squared residuals) of the Kalman filter on
    cpp_tests/data/carma_test.dat at the true parameters      -- cpp_tests/carma_unit_tests.cpp:387-502
  * 3-sigma posterior recovery for CAR(5), ZCARMA(5) and CARMA(5,4) on cpp_tests/data/{car5,zcar5,carma}_test.dat
    (many independent ensembles in one launch instead of one 75,000-iteration chain) -- :1378-1656
  * get_mle / choose_order: the GPU best-of-N fits against scipy L-BFGS-B on the CPU oracle FROM THE SAME STARTS
    (src/carmcmc/carma_pack.py:195-260), and a sharded AICc table against the single-process one, bitwise.
The fixtures are the committed copies in tests/golden/mcmc_fixtures.npz (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def C():
    import carma_pack_b200 as c
    if c._lib.device_count() < 1:
        pytest.fail("no CUDA device visible: GPU tests must run on the B200 box")
    return c


@pytest.fixture(scope="module")
def fx():
    return dict(np.load(os.path.join(GOLDEN, "mcmc_fixtures.npz")))


def true_ar_theta(fx, p=5):
    """log quadratic terms of the true AR polynomial (carma_unit_tests.cpp:1409-1424)."""
    th = np.empty(p)
    w, c = fx["qpo_width"], fx["qpo_cent"]
    for i in range(p // 2):
        re, im = -2 * np.pi * w[i], 2 * np.pi * c[i]
        th[2 * i], th[2 * i + 1] = np.log(re * re + im * im), np.log(-2 * re)
    if p % 2:
        th[p - 1] = np.log(2 * np.pi * w[p // 2])
    return th


def autocorr(x, maxlag):
    x = x - x.mean()
    den = np.dot(x, x)
    return np.array([np.dot(x[:-k], x[k:]) / den for k in range(1, maxlag + 1)])


def test_filter_residuals_are_white_on_reference_fixture(C, fx):
    """KalmanFilterp/Filter (carma_unit_tests.cpp:387-502): at the true CARMA(5,4) parameters the standardized
    one-step residuals pass Anderson-Darling (< 3.857) and the ACF tests on residuals and squared residuals."""
    from scipy.special import comb, ndtr
    from scipy.stats import chi2
    from carma_pack_b200 import synth
    t, y, e = fx["carma_t"], fx["carma_y"], fx["carma_yerr"]
    ny = t.size
    roots = synth.get_ar_roots(fx["qpo_width"], fx["qpo_cent"])
    kappa = float(fx["kappa"])
    ma = comb(4, np.arange(5)) / kappa ** np.arange(5)
    sigsqr = float(fx["sigmay"]) ** 2 / synth.carma_variance(1.0, roots, ma)
    s = C.Series(t, y, e)
    mean, var = s.filter(sigsqr, roots, ma)
    s.close()
    assert mean[0] == 0.0 and abs(var[0] - (2.3 ** 2 + e[0] ** 2)) < 1e-10
    sres = (y - mean) / np.sqrt(var)
    srt = np.sort(sres)
    cdf = ndtr(srt)
    i = np.arange(1, ny + 1)
    ad = -ny - np.sum((2.0 * i - 1) / ny * (np.log(cdf) + np.log(1.0 - cdf[::-1])))
    assert ad < 3.857, ad
    bound = 1.96 / np.sqrt(ny)
    for series in (sres, sres ** 2):
        ac = autocorr(series, 100)
        assert np.sum(np.abs(ac) > bound) < 11
        assert chi2.cdf(np.max(ac ** 2) * ny, 1) ** 100 < 0.99


def _pooled_z(samples, truth):
    """z-score of the truth against the pooled posterior (all ensembles' coolest chains)."""
    flat = samples.reshape(-1, samples.shape[-1])
    return (flat.mean(axis=0) - truth) / flat.std(axis=0, ddof=1), flat


@pytest.mark.parametrize("name", ["car5", "zcar5", "carma"])
def test_posterior_recovers_truth_within_3_sigma_on_reference_fixtures(C, fx, name):
    """./CAR5, ./ZCAR5, CARMA/mcmc_sampler (carma_unit_tests.cpp:1378-1656): every true parameter within 3 sigma of the
    marginal posterior mean.  128 independent ensembles (10 or 13 temperatures) run in ONE launch; their coolest
    chains are pooled after an adaptive burn-in."""
    from scipy.special import comb
    t, y, e = fx[name + "_t"], fx[name + "_y"], fx[name + "_yerr"]
    s = C.Series(t, y, e)
    prior = s.default_prior(population_var=True)
    ar = true_ar_theta(fx)
    if name == "car5":
        kind, p, q, ntemps = C.KIND_CARP, 5, 0, 10
        truth = np.concatenate([[np.log(2.3), 1.0, 0.0], ar])
    elif name == "zcar5":
        # the data are a ZCARMA(5) process with kappa = 0.5; fitted with the true model (free kappa)
        kind, p, q, ntemps = C.KIND_ZCARMA, 5, 0, 10
        kn = (0.5 - prior.kappa_low) / (prior.kappa_high - prior.kappa_low)
        truth = np.concatenate([[np.log(2.3), 1.0, 0.0], ar, [np.log(kn / (1.0 - kn))]])
    else:
        kind, p, q, ntemps = C.KIND_CARMA, 5, 4, 13
        truth = np.concatenate([[np.log(2.3), 1.0, 0.0], ar])
    res = s.pt_run(kind, p, q, nsamples=400, burnin=6000, thin=5, ntemps=ntemps, n_ensembles=128, seed=20260 + len(name),
                   prior=prior)
    samples = res["samples"].copy()
    assert np.all(np.isfinite(res["logposts"]))
    samples[..., 0] = np.log(samples[..., 0])       # the reference compares log(sigma_y)
    z, flat = _pooled_z(samples[..., :truth.size], truth)
    assert np.all(np.abs(z) < 3.0), (name, z)
    if name == "carma":
        # MA coefficients of every sample (ExtractMA) against the truth C(4,i)/kappa^i (carma_unit_tests.cpp:1601-1604, 1650-1656)
        from carma_pack_b200.carma_pack import CarmaSample
        th = res["samples"].reshape(-1, 12)
        cs = CarmaSample.__new__(CarmaSample)
        cs.p, cs.q, cs._samples = 5, 4, {}
        cs._ma_coefs(th)
        ma = cs._samples["ma_coefs"][:, 1:]
        ma_true = (comb(4, np.arange(5)) / 0.5 ** np.arange(5))[1:]
        zma = (ma.mean(axis=0) - ma_true) / ma.std(axis=0, ddof=1)
        assert np.all(np.abs(zma) < 3.0), zma
    # cold-chain acceptance of an adapted RAM sampler sits near its 0.25 target
    acc = res["accept_rates"][:, 0]
    assert 0.1 < np.median(acc) < 0.5, np.median(acc)
    s.close()


# ------------------------------------------------------------------------------------------------
# get_mle / choose_order parity (BASELINE config 4)
# ------------------------------------------------------------------------------------------------
def _scipy_fit(args):
    from scipy.optimize import minimize
    from oracle import oracle as O
    okind, p, q, t, y, e, x0, lo, hi, ign = args
    pr = O.default_prior(t, y)

    def nll(th):
        v = -O.logdensity(okind, p, q, t, y, e, th[None, :], prior=pr, ignore_prior=ign, fast=True)[0]
        return v if np.isfinite(v) else 1e300

    bounds = [(None if not np.isfinite(a) else a, None if not np.isfinite(b) else b) for a, b in zip(lo, hi)]
    r = minimize(nll, x0, method="L-BFGS-B", bounds=bounds)
    return float(r.fun)


def test_get_mle_best_of_n_matches_scipy_lbfgsb_on_the_oracle_from_the_same_starts(C):
    """For six (p,q) models spanning p = 2..7 on the ny = 500 config-4 series: the GPU fits (carma_mle_batch) and
    the reference's CPU path -- scipy.optimize.minimize(L-BFGS-B, finite-difference gradient) on the CPU oracle's
    -LogDensity with SetMLE(True), carma_pack.py:195-260 -- start from the SAME N points.  The GPU best-of-N must
    not be worse than the CPU best-of-N by more than 1e-3 (it is usually equal to ~1e-6 or slightly better)."""
    import multiprocessing as mp
    from carma_pack_b200 import synth
    from oracle import oracle as O
    O.build()
    t, y, e = synth.readme_series(500, 500)
    model = C.CarmaModel(t, y, e)
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 2)
    report = []
    with mp.get_context("fork").Pool(max(1, min(ncpu, 32))) as pool:
        for (p, q), n in (((2, 0), 100), ((3, 1), 100), ((4, 2), 100), ((5, 1), 64), ((6, 3), 48), ((7, 4), 48)):
            kind, x0, lo, hi, prior, flags = model.mle_starts(p, q, n, seed=500 + 31 * p + q)
            xg, fg, nit, nfev = model.series.mle_batch(kind, p, q, x0, lo, hi, prior=prior, flags=flags)
            okind = O.KIND_CARMA if q > 0 else O.KIND_CARP
            fc = np.array(pool.map(_scipy_fit, [(okind, p, q, t, y, e, x0[i], lo, hi, True) for i in range(n)]))
            # the GPU optimum is a genuine value of the objective: the oracle agrees at the GPU's theta-hat
            chk = -O.logdensity(okind, p, q, t, y, e, xg[np.argmin(fg)][None, :], prior=O.default_prior(t, y), ignore_prior=True)[0]
            assert abs(chk - fg.min()) <= 1e-7 * max(1.0, abs(chk)), (p, q, chk, fg.min())
            report.append((p, q, n, float(fg.min()), float(fc.min()), float(np.median(fg)), float(np.median(fc))))
            assert fg.min() <= fc.min() + 1e-3, report[-1]
    print("get_mle parity (p, q, N, GPU best, CPU best, GPU median, CPU median):")
    for r in report:
        print("   ", r)


def test_choose_order_table_does_not_depend_on_seed_beyond_a_small_tolerance(C):
    """Best-of-100 per model from two different sets of random starts reach the same AICc for the low orders (whose
    likelihood has one dominant mode) -- the round-1 tables differed by ~4 for some models because starts stalled."""
    from carma_pack_b200 import synth
    t, y, e = synth.readme_series(500, 500)
    model = C.CarmaModel(t, y, e)
    pq = [(1, 0), (2, 0), (2, 1), (3, 0), (3, 1), (3, 2), (4, 1)]
    _, _, a1 = model.choose_order(4, pqlist=pq, ntrials=100, seed=11, verbose=False)
    _, _, a2 = model.choose_order(4, pqlist=pq, ntrials=100, seed=977, verbose=False)
    assert np.allclose(a1, a2, atol=0.05), (a1, a2)


def _sharded_worker(rank, world, port, ndev, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import carma_pack_b200 as C
        from carma_pack_b200 import synth
        t, y, e = synth.readme_series(300, 300)
        model = C.CarmaModel(t, y, e, device=rank % ndev)
        mle, pqlist, aicc = model.choose_order(4, ntrials=24, seed=5, verbose=False, dist=dist)
        q.put((rank, list(aicc), np.asarray(mle.x).tolist(), (model.p, model.q)))
    finally:
        dist.destroy_process_group()


def test_sharded_choose_order_equals_single_process_bitwise(C):
    """Two ranks (two GPUs when the box has them, otherwise both on cuda:0) fit disjoint shares of the (model, start)
    grid with the REAL GPU optimiser; the gathered AICc table and theta-hat equal the single-process ones bit for
    bit: start j of model k depends only on (seed, k, j), and a start's iterates do not depend on its batch mates."""
    import socket
    import torch.multiprocessing as mp
    from carma_pack_b200 import synth
    t, y, e = synth.readme_series(300, 300)
    model = C.CarmaModel(t, y, e)
    mle1, pq1, aicc1 = model.choose_order(4, ntrials=24, seed=5, verbose=False)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ndev = C._lib.device_count()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, ndev, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(2)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for r in res:
        assert r[1] == list(aicc1), (r[1], list(aicc1))
        assert r[2] == np.asarray(mle1.x).tolist() and r[3] == (model.p, model.q)
