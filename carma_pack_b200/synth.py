"""Synthetic CARMA light curves and parameter batches (numpy only).

Re-implements the public helpers of the reference's Python layer that generate data
(src/carmcmc/carma_pack.py: get_ar_roots 1038-1059, carma_variance 1084-1123, car1_process
1126-1146, carma_process 1148-1259, power_spectrum 1062-1081) and the README recipe
(README.md:23-46) used by every BASELINE config.  Seeds use numpy.random.default_rng.
"""
import numpy as np


def get_ar_roots(qpo_width, qpo_centroid):
    """Roots of the AR characteristic polynomial from Lorentzian widths/centroids
    (carma_pack.py:1038-1059): each centroid > 0 contributes a conjugate pair."""
    ar_roots = []
    for i in range(len(qpo_centroid)):
        ar_roots.append(qpo_width[i] + 1j * qpo_centroid[i])
        if qpo_centroid[i] > 1e-10:
            ar_roots.append(np.conjugate(ar_roots[-1]))
    if len(qpo_width) - len(qpo_centroid) == 1:
        ar_roots.append(qpo_width[-1] + 1j * 0.0)
    return -2.0 * np.pi * np.array(ar_roots)


def power_spectrum(freq, sigma, ar_coef, ma_coefs=(1.0,)):
    """PSD of a CARMA(p,q) process (carma_pack.py:1062-1081)."""
    ma_coefs = np.asarray(ma_coefs, dtype=float)
    s = 2.0 * np.pi * 1j * np.asarray(freq)
    ma_poly = np.polyval(ma_coefs[::-1], s)
    ar_poly = np.polyval(ar_coef, s)
    return sigma ** 2 * np.abs(ma_poly) ** 2 / np.abs(ar_poly) ** 2


def carma_variance(sigsqr, ar_roots, ma_coefs=(1.0,), lag=0.0):
    """Autocovariance of a CARMA(p,q) process at `lag` (carma_pack.py:1084-1123)."""
    ar_roots = np.asarray(ar_roots, dtype=complex)
    p = ar_roots.size
    ma = np.zeros(p)
    ma[:len(ma_coefs)] = ma_coefs
    total = 0.0 + 0.0j
    powers = np.arange(p)
    for k in range(p):
        rk = ar_roots[k]
        others = np.delete(ar_roots, k)
        denom = -2.0 * rk.real * np.prod((others - rk) * (np.conjugate(others) + rk))
        num = np.sum(ma * rk ** powers) * np.sum(ma * (-rk) ** powers) * np.exp(rk * abs(lag))
        total += num / denom
    return sigsqr * total.real


def car1_process(time, sigsqr, tau, rng=None):
    """CAR(1)/OU process sampled at `time` (carma_pack.py:1126-1146)."""
    rng = np.random.default_rng() if rng is None else rng
    time = np.asarray(time, dtype=float)
    marginal_var = sigsqr * tau / 2.0
    y = np.zeros(time.size)
    z = rng.standard_normal(time.size)
    y[0] = np.sqrt(marginal_var) * z[0]
    for i in range(1, time.size):
        rho = np.exp(-(time[i] - time[i - 1]) / tau)
        y[i] = rho * y[i - 1] + np.sqrt(marginal_var * (1.0 - rho ** 2)) * z[i]
    return y


def carma_process(time, sigsqr, ar_roots, ma_coefs=(1.0,), rng=None):
    """Exact simulation of a CARMA(p,q) process at `time` (carma_pack.py:1148-1259): the rotated
    state-space filter is run without measurement noise and each value is drawn from its one-step
    predictive distribution (innovations form)."""
    rng = np.random.default_rng() if rng is None else rng
    time = np.sort(np.asarray(time, dtype=float))
    ar_roots = np.asarray(ar_roots, dtype=complex)
    p = ar_roots.size
    if p == 1:
        return car1_process(time, sigsqr, -1.0 / ar_roots[0].real, rng)
    ma = np.zeros(p)
    ma[:len(ma_coefs)] = ma_coefs
    E = np.vander(ar_roots, p, increasing=True).T  # E[i,k] = w_k^i
    R = np.zeros(p, dtype=complex)
    R[-1] = 1.0
    J = np.linalg.solve(E, R)
    b = ma @ E
    V = -sigsqr * np.outer(J, np.conjugate(J)) / (ar_roots[:, None] + np.conjugate(ar_roots)[None, :])
    P = V.copy()
    x = np.zeros(p, dtype=complex)
    bh = np.conjugate(b)
    z = rng.standard_normal(time.size)
    y = np.empty(time.size)
    mean, var = 0.0, float(np.real(b @ P @ bh))
    y[0] = mean + np.sqrt(var) * z[0]
    innov = y[0]
    for i in range(1, time.size):
        K = P @ bh / var
        x = x + innov * K
        P = P - var * np.outer(K, np.conjugate(K))
        rho = np.exp(ar_roots * (time[i] - time[i - 1]))
        x = rho * x
        P = np.outer(rho, np.conjugate(rho)) * (P - V) + V
        mean = float(np.real(b @ x))
        var = float(np.real(b @ P @ bh))
        y[i] = mean + np.sqrt(var) * z[i]
        innov = y[i] - mean
    return y


# ---- README recipe (README.md:23-46; src/paper/carma_paper.py:209-237) -------------------------
README = dict(sigmay=2.3, mu=17.0, p=5,
              qpo_width=np.array([1.0 / 100.0, 1.0 / 300.0, 1.0 / 200.0]),
              qpo_cent=np.array([1.0 / 5.0, 1.0 / 25.0]),
              ma_coefs=np.array([1.0, 4.5, 1.25, 0.0, 0.0]))


def readme_truth():
    ar_roots = get_ar_roots(README["qpo_width"], README["qpo_cent"])
    sigsqr = README["sigmay"] ** 2 / carma_variance(1.0, ar_roots, README["ma_coefs"])
    return ar_roots, README["ma_coefs"].copy(), sigsqr


def readme_times(ny, rng):
    """dt ~ U(1,3); three seasons of ny//3 points separated by 180 (README.md:36-41 for ny=270)."""
    s = ny // 3
    dt = rng.uniform(1.0, 3.0, ny)
    time = np.empty(ny)
    time[:s] = np.cumsum(dt[:s])
    time[s:2 * s] = 180 + time[s - 1] + np.cumsum(dt[s:2 * s])
    time[2 * s:] = 180 + time[2 * s - 1] + np.cumsum(dt[2 * s:])
    return time


def readme_series(ny=270, seed=270):
    """(time, y, ysig) of the README CARMA(5,q) light curve with `ny` points."""
    rng = np.random.default_rng(seed)
    ar_roots, ma, sigsqr = readme_truth()
    time = readme_times(ny, rng)
    y0 = README["mu"] + carma_process(time, sigsqr, ar_roots, ma, rng)
    ysig = np.ones(ny) * y0.std() / 5.0
    y = y0 + ysig * rng.standard_normal(ny)
    return time, y, ysig


def carma31_truth():
    """CARMA(3,1) used by the survey-scale config: one QPO pair + one low-frequency root."""
    ar_roots = get_ar_roots(np.array([1.0 / 50.0, 1.0 / 400.0]), np.array([1.0 / 20.0]))
    ma = np.array([1.0, 3.0, 0.0])
    sigsqr = 1.0 / carma_variance(1.0, ar_roots, ma)
    return ar_roots, ma, sigsqr


def carma31_theta(sigmay=1.0, mu=0.0):
    """theta (sigma_y, measerr scale, mu, log-quadratic AR terms, log-quadratic MA terms) of carma31_truth()."""
    ar_roots, ma, _ = carma31_truth()
    return np.array([sigmay, 1.0, mu] + list(roots_to_logquad(ar_roots)) + [np.log(1.0 / ma[1])])


def roots_to_logquad(ar_roots):
    """Inverse of CARp::ARRoots (carpack.cpp:137-172) for roots ordered as conjugate/real pairs
    followed by an optional single real root."""
    r = np.asarray(ar_roots, dtype=complex)
    p = r.size
    out = []
    for i in range(p // 2):
        a, b = r[2 * i], r[2 * i + 1]
        out += [np.log((a * b).real), np.log(-(a + b).real)]
    if p % 2:
        out.append(np.log(-r[-1].real))
    return np.array(out)


def readme_theta(q=3, far_root=50.0):
    """theta* of the README model written as CARMA(5,q).  The true MA polynomial
    1 + 4.5 s + 1.25 s^2 has q = 2; for q = 3 the extra MA root sits at -far_root."""
    ar_roots, ma, _ = readme_truth()
    th = [README["sigmay"], 1.0, README["mu"]] + list(roots_to_logquad(ar_roots))
    if q >= 2:
        th += [np.log(1.0 / 1.25), np.log(4.5 / 1.25)]
    if q == 1:
        th += [np.log(1.0 / 4.5)]
    if q == 3:
        th += [np.log(far_root)]
    if q > 3:
        raise ValueError("readme_theta supports q <= 3")
    return np.array(th)


def prior_draws(n, p, q, time, y, rng, kind="carma"):
    """Vectorised StartingValue draws (carpack.cpp:175-230, 268-311, 416-477) without the
    redraw-until-finite loop: rows can be outside the prior (they exercise the -inf path)."""
    time = np.asarray(time)
    ny = time.size
    dtm = np.diff(time)
    max_freq, min_freq = 1.0 / dtm.min(), 1.0 / (time.max() - time.min())
    nl = (p + 1) // 2
    lo, hi = np.log(min_freq), np.log(max_freq)
    cent = np.sort(np.exp(rng.uniform(lo, hi, (n, nl))), axis=1)[:, ::-1].copy()
    width = np.exp(rng.uniform(lo, hi, (n, nl)))
    if p % 2:
        cent[:, -1] = 0.0
        top = np.log(cent[:, -2]) if nl >= 2 else np.full(n, hi)
        width[:, -1] = np.exp(lo + (top - lo) * rng.uniform(size=n))
    th = np.empty((n, 3 + p + q))
    for i in range(p // 2):
        re, im = -2 * np.pi * width[:, i], 2 * np.pi * cent[:, i]
        th[:, 3 + 2 * i] = np.log(re * re + im * im)
        th[:, 4 + 2 * i] = np.log(-2 * re)
    if p % 2:
        th[:, 3 + p - 1] = np.log(2 * np.pi * width[:, -1])
    if q:
        th[:, 3 + p:] = np.abs(rng.standard_normal((n, q)))
    yvar = np.var(y, ddof=1) * (ny - 1) / rng.chisquare(ny - 1, n)
    th[:, 0] = np.sqrt(yvar)
    th[:, 2] = np.mean(y) + np.sqrt(yvar) / ny * rng.standard_normal(n)
    th[:, 1] = np.clip(50.0 / rng.chisquare(50, n), 0.51, 1.99)
    return th


def theta_batch(n, time, y, p=5, q=3, seed=0):
    """BASELINE config-2 batch: half = theta* + Sigma0^(1/2) t_8 perturbations with Sigma0 of
    carmcmc.cpp:127-136 (inflated by 25 so the batch is not a single point), half = prior draws."""
    rng = np.random.default_rng(seed)
    ny = len(y)
    d = 3 + p + q
    var = np.var(y)
    sd = np.full(d, 0.01)
    sd[0] = np.sqrt(2.0 * var * var / ny)
    sd[2] = np.sqrt(var / ny)
    n1 = n // 2
    t8 = rng.standard_t(8, size=(n1, d))
    th1 = readme_theta(q)[None, :] + 5.0 * sd[None, :] * t8 if p == 5 else None
    if th1 is None:
        th1 = prior_draws(n1, p, q, time, y, rng)
    th2 = prior_draws(n - n1, p, q, time, y, rng)
    return np.ascontiguousarray(np.vstack([th1, th2]))


def cauchy_times(ny, rng, dt_min=0.1, dt_max=1e3):
    """Irregular sampling of cpp_tests/generate_test_data.py:17: dt = 0.1 + |Cauchy|, truncated."""
    dt = np.minimum(dt_min + np.abs(rng.standard_cauchy(ny)), dt_max)
    return np.cumsum(dt)
