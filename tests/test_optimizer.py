"""The batched projected L-BFGS that replaces the reference's per-trial scipy L-BFGS-B fits
(src/carmcmc/carma_pack.py:195-252) -- host logic, checked on CPU against scipy."""
import numpy as np
from scipy.optimize import minimize

from carma_pack_b200 import batched_lbfgs, synth
from oracle import oracle as O


def test_batched_lbfgs_rosenbrock_box():
    def f(z):
        return (1 - z[:, 0]) ** 2 + 100.0 * (z[:, 1] - z[:, 0] ** 2) ** 2 + (z[:, 2] - 0.3) ** 2

    rng = np.random.default_rng(0)
    x0 = rng.uniform(-1.5, 1.5, (16, 3))
    lo = np.array([-2.0, -2.0, 0.5])   # third coordinate's optimum (0.3) is outside the box
    hi = np.array([2.0, 2.0, 2.0])
    x, fv, nit, nfev = batched_lbfgs(f, x0, lo, hi, maxiter=500)
    assert np.all(x >= lo - 1e-15) and np.all(x <= hi + 1e-15)
    assert np.allclose(x[:, 2], 0.5)
    good = np.abs(fv - 0.04) < 1e-3
    assert good.mean() > 0.8, fv
    assert np.allclose(x[good, :2], 1.0, atol=3e-2)


def test_batched_lbfgs_matches_scipy_on_car_loglik():
    """Same objective as _carma_loglik (carma_pack.py:255-260), CAR(2) on a short series, evaluated by
    the CPU oracle: the best of a batch reaches the L-BFGS-B optimum."""
    t, y, e = synth.readme_series(90, 3)
    pr = O.default_prior(t, y)

    def nll(th):
        return -O.logdensity(O.KIND_CARP, 2, 0, t, y, e, np.atleast_2d(th), prior=pr, ignore_prior=True)

    rng = np.random.default_rng(1)
    x0 = synth.prior_draws(12, 2, 0, t, y, rng)
    x0[:, 1] = 1.0
    ysig = y.std()
    lo = np.array([ysig / 10, 0.9, -np.inf, -12.0, -12.0])
    hi = np.array([10 * ysig, 1.1, np.inf, 3.0, 3.0])
    x0 = np.clip(x0, lo, hi)
    x, fv, nit, nfev = batched_lbfgs(nll, x0, lo, hi, maxiter=150)
    best = fv.min()
    ref = min(minimize(lambda th: float(nll(th)[0]), x0[k], method="L-BFGS-B",
                       bounds=list(zip(np.where(np.isfinite(lo), lo, None), np.where(np.isfinite(hi), hi, None)))).fun
              for k in range(4))
    assert best <= ref + 0.05, (best, ref)
    assert np.isfinite(fv).all()


# ---- the native optimiser core (carma_lbfgs_batch, csrc/mle.cu) against its numpy twin: host code, runs on CPU
def _native():
    from carma_pack_b200 import _lib
    return _lib.lbfgs_batch


def test_native_lbfgs_core_same_iterates_as_numpy_twin():
    """Same algorithm, so on a smooth objective both reach the same optima from the same starts with the same
    number of iterations (they differ only in summation order and in how many trial points a launch carries)."""
    def f(z):
        return (1 - z[:, 0]) ** 2 + 100.0 * (z[:, 1] - z[:, 0] ** 2) ** 2 + (z[:, 2] - 0.3) ** 2

    rng = np.random.default_rng(0)
    x0 = rng.uniform(-1.5, 1.5, (16, 3))
    lo = np.array([-2.0, -2.0, 0.5])
    hi = np.array([2.0, 2.0, 2.0])
    xa, fa, nita, nfeva = _native()(f, x0, lo, hi, maxiter=500)
    xb, fb, nitb, nfevb = batched_lbfgs(f, x0, lo, hi, maxiter=500)
    assert np.all(xa >= lo - 1e-15) and np.all(xa <= hi + 1e-15)
    assert np.allclose(xa[:, 2], 0.5)
    assert abs(nita - nitb) <= 2
    # finite-difference noise moves the late iterates of the twin and the core apart by rounding only
    assert np.allclose(fa, fb, rtol=1e-5, atol=1e-7), np.abs(fa - fb).max()
    assert np.allclose(xa, xb, atol=2e-3)
    assert nfeva > 0


def test_native_lbfgs_core_quadratic_with_active_bounds_and_failures():
    """Convex quadratic with a known box-constrained optimum; rows whose objective is never finite are returned
    untouched with f = 1e300; an objective that raises is reported as an error."""
    from carma_pack_b200 import CarmaError
    A = np.array([[3.0, 0.5, 0.0, 0.0], [0.5, 2.0, 0.3, 0.0], [0.0, 0.3, 1.0, 0.2], [0.0, 0.0, 0.2, 4.0]])
    c = np.array([1.0, -2.0, 0.5, 3.0])

    def f(z):
        v = 0.5 * np.einsum("ni,ij,nj->n", z - c, A, z - c)
        v[z[:, 0] > 50.0] = np.nan      # a region without finite values
        return v

    lo = np.array([-5.0, -5.0, 1.0, -np.inf])   # optimum of coordinate 2 (0.5) is below its lower bound
    hi = np.array([0.5, 5.0, 5.0, np.inf])      # optimum of coordinate 0 (1.0) is above its upper bound
    rng = np.random.default_rng(3)
    x0 = rng.uniform(-4, 4, (10, 4))
    x0[:, 0] = np.minimum(x0[:, 0], 0.5)
    x, fv, nit, nfev = _native()(f, x0, lo, hi, maxiter=200)
    # reference solution: scipy on the same box
    ref = minimize(lambda z: float(f(z[None, :])[0]), np.array([0.0, 0.0, 1.5, 0.0]), method="L-BFGS-B",
                   bounds=[(-5, 0.5), (-5, 5), (1, 5), (None, None)])
    assert np.allclose(fv, ref.fun, rtol=1e-6, atol=1e-8)
    assert np.allclose(x, ref.x[None, :], atol=2e-4)
    assert np.allclose(x[:, 0], 0.5) and np.allclose(x[:, 2], 1.0)
    # never-finite rows
    xbad = np.full((3, 4), 0.0)
    xbad[:, 0] = 60.0
    xo, fo, _, _ = _native()(f, xbad, np.full(4, -100.0), np.full(4, 100.0))
    assert np.all(fo >= 1e300) and np.array_equal(xo, xbad)

    def boom(z):
        raise RuntimeError("no")

    try:
        _native()(boom, x0, lo, hi)
        raise AssertionError("expected CarmaError")
    except CarmaError:
        pass


def test_native_lbfgs_core_matches_twin_on_car_loglik():
    """The real objective (CAR(2) negative log-likelihood from the CPU oracle) through both implementations."""
    t, y, e = synth.readme_series(90, 3)
    pr = O.default_prior(t, y)

    def nll(th):
        return -O.logdensity(O.KIND_CARP, 2, 0, t, y, e, np.atleast_2d(th), prior=pr, ignore_prior=True)

    rng = np.random.default_rng(1)
    x0 = synth.prior_draws(12, 2, 0, t, y, rng)
    x0[:, 1] = 1.0
    ysig = y.std()
    lo = np.array([ysig / 10, 0.9, -np.inf, -12.0, -12.0])
    hi = np.array([10 * ysig, 1.1, np.inf, 3.0, 3.0])
    x0 = np.clip(x0, lo, hi)
    xa, fa, nita, _ = _native()(nll, x0, lo, hi, maxiter=150)
    xb, fb, nitb, _ = batched_lbfgs(nll, x0, lo, hi, maxiter=150)
    assert np.isfinite(fa).all()
    assert abs(fa.min() - fb.min()) < 1e-4 * max(1.0, abs(fb.min()))
    assert np.mean(np.abs(fa - fb) < 1e-2 * np.maximum(1.0, np.abs(fb))) >= 0.75


def test_rows_do_not_depend_on_their_batch_mates():
    """Every start keeps its own L-BFGS history: a row whose (s, y) pair fails the curvature test must not be
    handed a zero pair because some OTHER row passed it (that collapsed H0 to 1e-8 I and froze the row far from a
    stationary point).  Non-convex objective with negative curvature near the origin; the result of a row is the
    same whether it runs alone or next to other starts -- in the numpy twin and in the native core."""
    def f(z):
        return np.sum(z ** 4 - z ** 2 + 0.1 * z, axis=1)

    lo = np.full(2, -10.0)
    hi = np.full(2, 10.0)
    a = np.array([[0.05, 0.02]])
    both = np.array([[0.05, 0.02], [2.0, 3.0]])
    fmin = 2 * min(v ** 4 - v ** 2 + 0.1 * v for v in np.linspace(-1.5, 1.5, 300001))
    for opt in (batched_lbfgs, _native()):
        xs, fs, _, _ = opt(f, a, lo, hi, maxiter=200)
        xb, fb, _, _ = opt(f, both, lo, hi, maxiter=200)
        assert np.array_equal(xs[0], xb[0]) and fs[0] == fb[0], (opt, xs, xb)
        assert fb[0] < -0.5, fb                     # a genuine local minimum, not the stall at f = -0.003
        assert fb.min() <= fmin + 1e-6 or fb[0] <= fmin + 0.15, (fb, fmin)
        # order inside the batch is irrelevant too
        xr, fr, _, _ = opt(f, both[::-1].copy(), lo, hi, maxiter=200)
        assert np.array_equal(xr[::-1], xb) and np.array_equal(fr[::-1], fb)


def test_gradient_next_to_an_infeasible_region_uses_the_other_side():
    """A forward-difference point without a finite value is retried with a backward difference, so a start that
    sits next to the edge of the support is not reported as stationary there."""
    def f(z):
        v = (z[:, 0] + 3.0) ** 2 + (z[:, 1] + 1.0) ** 2
        v[z[:, 0] > 1.0] = np.inf           # the density has no support beyond x0 = 1
        return v

    lo = np.full(2, -50.0)
    hi = np.full(2, 50.0)
    x0 = np.array([[1.0, 4.0], [1.0 - 5e-9, -3.0]])   # forward step (+1e-8) leaves the support for both rows
    for opt in (batched_lbfgs, _native()):
        x, fv, _, _ = opt(f, x0, lo, hi, maxiter=100)
        # with the component zeroed (old behaviour) x0 would have stayed at 1 and f at 16
        assert np.allclose(x, [[-3.0, -1.0]] * 2, atol=1e-4), (opt, x)
        assert np.all(fv < 1e-7), (opt, fv)
