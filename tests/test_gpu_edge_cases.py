"""Edge cases of the GPU path: extreme sizes, maximum orders, degenerate inputs, error behaviour."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def C():
    import carma_pack_b200 as c
    if c._lib.device_count() < 1:
        pytest.fail("no CUDA device visible")
    return c


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def test_tiny_and_odd_batch_sizes(C, O):
    from carma_pack_b200 import synth
    t, y, e = synth.readme_series(30, 1)
    for ny in (2, 3, 30):
        s = C.Series(t[:ny], y[:ny], e[:ny])
        pr = s.default_prior()
        opr = O.default_prior(t[:ny], y[:ny])
        for n in (1, 63, 64, 65, 1000):
            th = synth.prior_draws(n, 3, 1, t, y, np.random.default_rng(n))
            got = s.loglik(C.KIND_CARMA, 3, 1, th, prior=pr, flags=C.IGNORE_BOUNDS)
            want = O.logdensity(O.KIND_CARMA, 3, 1, t[:ny], y[:ny], e[:ny], th, prior=opr, ignore_prior=True)
            np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9)
        assert s.loglik(C.KIND_CARMA, 3, 1, np.empty((0, 7)), prior=pr).shape == (0,)
        s.close()


def test_max_order_p7_q6(C, O):
    from carma_pack_b200 import synth
    t, y, e = synth.readme_series(150, 2)
    s = C.Series(t, y, e)
    th = synth.prior_draws(200, 7, 6, t, y, np.random.default_rng(0))
    assert th.shape[1] == 16
    got = s.loglik(C.KIND_CARMA, 7, 6, th)
    want = O.logdensity(O.KIND_CARMA, 7, 6, t, y, e, th)
    fin = np.isfinite(want)
    assert np.array_equal(fin, np.isfinite(got)) and fin.mean() > 0.3
    rel = np.abs(got[fin] - want[fin]) / np.maximum(np.abs(want[fin]), 1)
    assert np.median(rel) < 1e-12 and np.quantile(rel, 0.9) < 1e-9
    res = s.pt_run(C.KIND_CARMA, 7, 6, 20, 30, ntemps=12, n_ensembles=2, seed=3)
    assert np.all(np.isfinite(res["logposts"]))
    relp = s.loglik(C.KIND_CARMA, 7, 6, res["samples"].reshape(-1, 16)).reshape(res["logposts"].shape)
    np.testing.assert_allclose(relp, res["logposts"], rtol=1e-8)
    s.close()


def test_nan_and_inf_parameters_are_values_not_errors(C, O):
    from carma_pack_b200 import synth
    t, y, e = synth.readme_series(60, 4)
    s = C.Series(t, y, e)
    th = np.tile(synth.readme_theta(3), (6, 1))
    th[1, 0] = np.nan
    th[2, 3] = np.inf
    th[3, 2] = -np.inf
    th[4, 5] = np.nan
    th[5, 1] = np.nan
    got = s.loglik(C.KIND_CARMA, 5, 3, th)
    want = O.logdensity(O.KIND_CARMA, 5, 3, t, y, e, th)
    assert np.isfinite(got[0]) and abs(got[0] - want[0]) < 1e-9 * abs(want[0])
    assert not np.isfinite(got[1:]).any() and not np.isfinite(want[1:]).any()
    # a chain fed NaN never accepts it (steps.cpp:41-46): the sampler keeps running
    res = s.pt_run(C.KIND_CARMA, 5, 3, 10, 10, ntemps=3, n_ensembles=1, seed=1)
    assert np.all(np.isfinite(res["samples"]))
    s.close()


def test_tiny_measurement_error_and_large_gaps(C, O):
    """Gaps of 5e4 and 3e6 time units (exp underflows to exactly 0: the state forgets everything) and
    measurement errors 100x below the signal.  (Exactly zero errors make the filter variance collapse
    to rounding noise of either sign, where no two implementations agree on finiteness.)"""
    from carma_pack_b200 import synth
    rng = np.random.default_rng(8)
    t = np.cumsum(np.concatenate([rng.uniform(0.5, 1.5, 40), [5e4], rng.uniform(0.5, 1.5, 40), [3e6], rng.uniform(1, 2, 10)]))
    ar, ma, s2 = synth.carma31_truth()
    y = synth.carma_process(t, s2, ar, ma, rng)
    e = np.full(t.size, 1e-2)
    y = y + e * rng.standard_normal(t.size)
    s = C.Series(t, y, e)
    th0 = np.array([1.0, 1.0, 0.0] + list(synth.roots_to_logquad(ar)) + [np.log(1.0 / 3.0)])
    th = th0[None, :] + 0.2 * rng.standard_normal((64, 7))
    th[:, 1] = np.clip(th[:, 1], 0.6, 1.9)
    got = s.loglik(C.KIND_CARMA, 3, 1, th, flags=C.IGNORE_BOUNDS)
    want = O.logdensity(O.KIND_CARMA, 3, 1, t, y, e, th, ignore_prior=True)
    assert np.all(np.isfinite(want)) and np.all(np.isfinite(got))
    np.testing.assert_allclose(got, want, rtol=1e-9)
    # the scan kernel sees the same gaps
    got2 = s.loglik_scan(C.KIND_CARMA, 3, 1, th[:4], flags=C.IGNORE_BOUNDS, chunk=8)
    np.testing.assert_allclose(got2, want[:4], rtol=1e-9)
    s.close()


@pytest.mark.parametrize("ntemps", [1, 2, 33, 64])
def test_pt_ladder_sizes(C, ntemps):
    from carma_pack_b200 import synth
    t, y, e = synth.readme_series(60, 5)
    s = C.Series(t, y, e)
    res = s.pt_run(C.KIND_CARP, 2, 0, 15, 15, ntemps=ntemps, n_ensembles=3, seed=ntemps)
    assert res["samples"].shape == (3, 15, 5) and np.all(np.isfinite(res["logposts"]))
    assert res["accept_rates"].shape == (3, ntemps)
    s.close()


def test_error_paths(C):
    from carma_pack_b200 import synth
    t, y, e = synth.readme_series(60, 6)
    s = C.Series(t, y, e)
    with pytest.raises(C.CarmaError):
        s.loglik(C.KIND_CARMA, 8, 0, np.zeros((1, 11)))          # p out of range
    with pytest.raises(C.CarmaError):
        s.loglik(C.KIND_CARMA, 3, 3, np.zeros((1, 9)))           # q >= p
    with pytest.raises(ValueError):
        s.loglik(C.KIND_CARMA, 3, 1, np.zeros((1, 6)))           # wrong theta width
    with pytest.raises(C.CarmaError):
        s.pt_run(C.KIND_CARP, 2, 0, 5, 5, ntemps=65)              # ladder too long for one block
    with pytest.raises(C.CarmaError):
        s.pt_run(C.KIND_CARP, 2, 0, 5, 5, dof=7)                  # odd Student-t dof is not supported
    s.close()
    # the series object is still usable after errors
    s2 = C.Series(t, y, e)
    assert np.isfinite(s2.loglik(C.KIND_CARMA, 5, 3, synth.readme_theta(3))[0])
    s2.close()


def test_pt_on_series_too_long_for_shared_memory(C, O):
    """ny = 12,000 does not fit the resident-series staging: the sampler reads the series from global memory
    and must give exactly the chain the oracle's sequential sampler produces from the same streams."""
    rng = np.random.default_rng(12)
    tl = np.cumsum(rng.uniform(0.5, 1.5, 12000))
    yl = np.sin(tl / 30.0) + 0.3 * rng.standard_normal(tl.size)
    el = np.full(tl.size, 0.3)
    sl = C.Series(tl, yl, el)
    pr = sl.default_prior()
    res = sl.pt_run(C.KIND_CARP, 2, 0, 6, 6, ntemps=3, n_ensembles=2, seed=5, prior=pr)
    assert np.all(np.isfinite(res["logposts"]))
    relp = sl.loglik(C.KIND_CARP, 2, 0, res["samples"].reshape(-1, 5), prior=pr).reshape(res["logposts"].shape)
    np.testing.assert_allclose(relp, res["logposts"], rtol=1e-9)
    ores = O.pt_run(O.KIND_CARP, 2, 0, tl, yl, el, 6, 6, ntemps=3, seed=5, prior=O.default_prior(tl, yl))
    np.testing.assert_allclose(res["samples"][0], ores["samples"], rtol=1e-7, atol=1e-9)
    sl.close()


def test_reported_cuda_error_does_not_leak_into_later_calls(C):
    """An allocation failure is reported once (CarmaError) and leaves the library usable: the stale CUDA error
    must not resurface in the launch check of a later, unrelated call."""
    from carma_pack_b200 import synth
    th = synth.carma31_theta()
    with pytest.raises(C.CarmaError):
        C.MultiSeries.simulate(40_000_000, 1000, C.KIND_CARMA, 3, 1, th)   # 960 GB: cudaMalloc fails
    t, y, e = synth.readme_series(120, 3)
    s = C.Series(t, y, e)
    got = s.loglik_scan(C.KIND_CARMA, 3, 1, th[None, :], flags=C.IGNORE_BOUNDS)
    seq = s.loglik(C.KIND_CARMA, 3, 1, th[None, :], flags=C.IGNORE_BOUNDS)
    assert np.isfinite(got[0]) and abs(got[0] - seq[0]) <= 1e-9 * abs(seq[0])
    s.close()
