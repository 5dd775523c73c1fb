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
