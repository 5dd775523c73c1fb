#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE's own Python code.

Run only in the build container (needs /root/reference, read-only).  The reference package
cannot be imported (no compiled `_carmcmc`, no matplotlib), so the pure-numpy definitions are
lifted out of src/carmcmc/carma_pack.py with `ast` and exec'd unchanged:

  * KalmanFilterDeprecated (carma_pack.py:1264-1488)  reset/update/filter/predict
  * carma_variance (1084-1123), get_ar_roots (1038-1059), carma_process (1148-1259)
  * CarmaSample._ar_roots/_ma_coefs/_sigma_noise (439-546): theta -> (roots, MA coefs, sigma)

Nothing from the reference is copied into this repository: only numbers produced by running it.
The fixtures hold the inputs (series + parameters) and the reference outputs.
"""
import ast
import os
import sys
import warnings

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

warnings.filterwarnings("ignore")


def load_reference_namespace():
    src = open(os.path.join(REF, "src/carmcmc/carma_pack.py")).read()
    tree = ast.parse(src)
    wanted_funcs = {"get_ar_roots", "carma_variance", "carma_process", "car1_process", "power_spectrum"}
    wanted_classes = {"KalmanFilterDeprecated"}
    body = []
    methods = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in wanted_funcs:
            body.append(node)
        elif isinstance(node, ast.ClassDef) and node.name in wanted_classes:
            body.append(node)
        elif isinstance(node, ast.ClassDef) and node.name == "CarmaSample":
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name in ("_ar_roots", "_ma_coefs", "_sigma_noise"):
                    methods[sub.name] = sub
    shim = ast.ClassDef(name="ThetaShim", bases=[], keywords=[], body=list(methods.values()), decorator_list=[])
    try:
        shim.type_params = []
    except Exception:
        pass
    body.append(shim)
    mod = ast.Module(body=body, type_ignores=[])
    ast.fix_missing_locations(mod)
    from scipy.linalg import solve
    ns = {"np": np, "solve": solve}
    exec(compile(mod, "<reference carma_pack.py>", "exec"), ns)
    return ns


NS = load_reference_namespace()


def theta_to_params(theta, p, q):
    """Reference-Python theta -> (ar_roots, ma_coefs (len p), sigsqr) via CarmaSample methods."""
    theta = np.atleast_2d(np.asarray(theta, dtype=float))
    shim = NS["ThetaShim"]()
    shim.p = p
    shim.q = q
    # samplers.py / carma_pack.py:  var = sigma_y^2, quad_coefs = exp(theta[3:3+p])
    shim._samples = {"var": theta[:, 0] ** 2, "quad_coefs": np.exp(theta[:, 3:3 + p])}
    shim._ar_roots()
    shim._ma_coefs(theta)
    shim._sigma_noise()
    roots = shim._samples["ar_roots"]
    ma = shim._samples["ma_coefs"]
    ma_full = np.zeros((theta.shape[0], p))
    ma_full[:, :ma.shape[1]] = ma
    sigsqr = shim._samples["sigma"] ** 2
    return roots, ma_full, sigsqr


def ref_filter(t, ycent, yvar, sigsqr, roots, ma):
    kf = NS["KalmanFilterDeprecated"](t, ycent, yvar, sigsqr, np.asarray(roots, dtype=complex), ma_coefs=list(ma))
    mean, var = kf.filter()
    return np.array(mean, dtype=float), np.array(var, dtype=float)


def ref_logpost(theta, p, q, t, y, ysig, zcarma_kappa_bounds=None):
    """carpack.hpp:131-176 evaluated with the reference's numpy filter (no bounds check)."""
    theta = np.asarray(theta, dtype=float)
    if zcarma_kappa_bounds is None:
        roots, ma, sigsqr = theta_to_params(theta, p, q)
        roots, ma, sigsqr = roots[0], ma[0], float(sigsqr[0])
    else:
        roots, _, _ = theta_to_params(theta[:3 + p], p, 0)
        roots = roots[0]
        klo, khi = zcarma_kappa_bounds
        x = theta[3 + p]
        kappa = (khi - klo) * (np.exp(x) / (1.0 + np.exp(x))) + klo
        from scipy.special import comb
        ma = np.array([comb(p - 1, i, exact=True) / kappa ** i for i in range(p)])
        sigsqr = theta[0] ** 2 / NS["carma_variance"](1.0, roots, ma_coefs=ma)
    scale, mu = theta[1], theta[2]
    mean, var = ref_filter(t, y - mu, scale * ysig ** 2, sigsqr, roots, ma)
    ll = np.sum(-0.5 * np.log(var) - 0.5 * (y - mu - mean) ** 2 / var)
    prior = -0.5 * 50.0 / scale - (1.0 + 50.0 / 2.0) * np.log(scale)
    if zcarma_kappa_bounds is not None:
        prior += -theta[3 + p] - 2.0 * np.log(1.0 + np.exp(-theta[3 + p]))
    return ll + prior, ll, prior


def make_kelly_fixture():
    """cpp_tests/data/carma_test.dat with the parameters of TEST_CASE KalmanFilterp/Filter
    (carma_unit_tests.cpp:387-444) -- SURVEY 8(c) golden vector."""
    data = np.loadtxt(os.path.join(REF, "cpp_tests/data/carma_test.dat"))
    t, y, yerr = data[:, 0].copy(), data[:, 1].copy(), data[:, 2].copy()
    p = 5
    widths = np.array([0.01, 0.01, 0.002])
    cents = np.array([0.2, 0.02])
    roots = []
    for i in range(2):
        roots += [complex(-2 * np.pi * widths[i], 2 * np.pi * cents[i]), complex(-2 * np.pi * widths[i], -2 * np.pi * cents[i])]
    roots.append(complex(-2 * np.pi * widths[2], 0.0))
    roots = np.array(roots)
    from scipy.special import comb
    kappa = 0.5
    ma = np.array([comb(p - 1, i, exact=True) / kappa ** i for i in range(p)])
    sigmay = 2.3
    sigsqr = sigmay ** 2 / NS["carma_variance"](1.0, roots, ma_coefs=ma)
    mean, var = ref_filter(t, y, yerr ** 2, sigsqr, roots, ma)
    out = dict(t=t, y=y, yerr=yerr, roots=roots, ma=ma, sigsqr=sigsqr, mean=mean, var=var)
    out["loglik_270"] = np.sum(-0.5 * np.log(var[:270]) - 0.5 * (y[:270] - mean[:270]) ** 2 / var[:270])
    # NB: prefix sums of the full-series filter (the filter is causal)
    out["loglik_500"] = np.sum(-0.5 * np.log(var[:500]) - 0.5 * (y[:500] - mean[:500]) ** 2 / var[:500])
    out["loglik_1000"] = np.sum(-0.5 * np.log(var) - 0.5 * (y - mean) ** 2 / var)
    # with measerr_scale = 1.3 and mu = 0.25
    scale, mu = 1.3, 0.25
    m2, v2 = ref_filter(t, y - mu, scale * yerr ** 2, sigsqr, roots, ma)
    out["scaled_loglik"] = np.sum(-0.5 * np.log(v2) - 0.5 * (y - mu - m2) ** 2 / v2)
    out["scaled_logprior"] = -0.5 * 50.0 / scale - 26.0 * np.log(scale)
    # predictions (reference numpy predict supports interpolation and forecasting)
    span = t.max() - t.min()
    tq = np.array([166.0, t[-1] + 0.05 * span, 0.5 * (t[10] + t[11]), t[-1] + 1e-3 * span, 4000.123])
    pm, pv = [], []
    for tt in tq:
        kf = NS["KalmanFilterDeprecated"](t, y, yerr ** 2, sigsqr, roots, ma_coefs=list(ma))
        m_, v_ = kf.predict(tt)
        pm.append(float(m_))
        pv.append(float(v_))
    out["predict_t"] = tq
    out["predict_mean"] = np.array(pm)
    out["predict_var"] = np.array(pv)
    # ZCAR/variance known answer (carma_unit_tests.cpp:1268-1317): kappa = 0.7, sigma = 2.3
    ma07 = np.array([comb(p - 1, i, exact=True) / 0.7 ** i for i in range(p)])
    out["variance_kappa07"] = NS["carma_variance"](2.3 ** 2, roots, ma_coefs=ma07)
    out["variance_kappa07_published"] = 223003.230567
    # autocovariance at a few lags
    lags = np.array([0.0, 0.5, 3.0, 40.0])
    out["acov_lags"] = lags
    out["acov"] = np.array([NS["carma_variance"](sigsqr, roots, ma_coefs=ma, lag=l) for l in lags])
    np.savez(os.path.join(HERE, "kelly_carma_test.npz"), **out)
    print("kelly fixture: sigsqr=%.13e var0=%.10f ll1000=%.12f" % (sigsqr, var[0], out["loglik_1000"]))
    return out


def readme_series(ny, seed):
    """README.md:23-46 recipe (same as src/paper/carma_paper.py:209-237) using the reference's
    carma_process; seasons of ny/3 separated by 180."""
    np.random.seed(seed)
    sigmay, p, mu = 2.3, 5, 17.0
    qpo_width = np.array([1.0 / 100.0, 1.0 / 300.0, 1.0 / 200.0])
    qpo_cent = np.array([1.0 / 5.0, 1.0 / 25.0])
    ar_roots = NS["get_ar_roots"](qpo_width, qpo_cent)
    ma_coefs = np.zeros(p)
    ma_coefs[0], ma_coefs[1], ma_coefs[2] = 1.0, 4.5, 1.25
    sigsqr = sigmay ** 2 / NS["carma_variance"](1.0, ar_roots, ma_coefs=ma_coefs)
    s = ny // 3
    time = np.empty(ny)
    dt = np.random.uniform(1.0, 3.0, ny)
    time[:s] = np.cumsum(dt[:s])
    time[s:2 * s] = 180 + time[s - 1] + np.cumsum(dt[s:2 * s])
    time[2 * s:] = 180 + time[2 * s - 1] + np.cumsum(dt[2 * s:])
    y0 = mu + NS["carma_process"](time, sigsqr, ar_roots, ma_coefs=ma_coefs)
    ysig = np.ones(ny) * y0.std() / 5.0
    y = y0 + ysig * np.random.standard_normal(ny)
    return time, y, ysig, ar_roots, ma_coefs, sigsqr


def true_theta_53(ar_roots, ma_coefs, sigmay=2.3, mu=17.0):
    """theta of the README CARMA(5,3)-parameterised truth: ma_coefs [1,4.5,1.25] is a q=2
    polynomial; as a q=3 model the third MA root is sent far away (large log linear term)."""
    th = [sigmay, 1.0, mu]
    # AR: pairs (conjugates adjacent); quad terms: |w|^2 and -2 Re w
    r = ar_roots
    for i in range(2):
        w = r[2 * i]
        th += [np.log(abs(w) ** 2), np.log(-2.0 * w.real)]
    th.append(np.log(-r[4].real))
    return np.array(th)


def make_loglik_fixture():
    """CARMA(p,q) LogDensity over theta batches on a README-style ny=270 series (config C1/C2 shape)
    and on a short ny=60 series for the other (p,q)."""
    t, y, ysig, ar_roots, ma_true, sigsqr = readme_series(270, 270)
    rng = np.random.default_rng(12345)
    base = true_theta_53(ar_roots, ma_true)
    cases = {}
    # ---- CARMA(5,3): MA params = log quad terms of 3 roots
    # MA polynomial 1 + 4.5 s + 1.25 s^2 = 1.25 (s^2 + 3.6 s + 0.8): roots real; q=3 adds root at -50
    ma_log = np.array([np.log(0.8), np.log(3.6), np.log(50.0)])
    th0 = np.concatenate([base, ma_log])
    thetas = [th0]
    for k in range(23):
        th = th0.copy()
        th[0] *= np.exp(0.2 * rng.standard_normal())
        th[1] = np.clip(1.0 + 0.2 * rng.standard_normal(), 0.55, 1.9)
        th[2] += 0.5 * rng.standard_normal()
        th[3:] += 0.4 * rng.standard_normal(th.size - 3)
        thetas.append(th)
    # force a real AR pair (discriminant > 0) and a complex MA pair in some rows
    th = th0.copy(); th[3] = np.log(0.01); th[4] = np.log(0.5); thetas.append(th)
    th = th0.copy(); th[8] = np.log(2.0); th[9] = np.log(0.5); thetas.append(th)
    thetas = np.array(thetas)
    lp = np.array([ref_logpost(th, 5, 3, t, y, ysig) for th in thetas])
    cases["c53"] = dict(p=5, q=3, theta=thetas, logpost=lp[:, 0], loglik=lp[:, 1], logprior=lp[:, 2])
    print("c53 logpost[0..3] =", lp[:4, 0])
    # ---- other (p,q) on a shorter series
    ts, ys, es = t[:60].copy(), y[:60].copy(), ysig[:60].copy()
    for (p, q) in [(2, 0), (2, 1), (3, 0), (3, 2), (4, 1), (4, 3), (6, 0), (6, 4), (7, 0), (7, 6), (5, 0)]:
        rows = []
        for k in range(6):
            th = [2.3 * np.exp(0.2 * rng.standard_normal()), np.clip(1 + 0.2 * rng.standard_normal(), 0.55, 1.9),
                  17.0 + 0.5 * rng.standard_normal()]
            # lorentzian-style AR terms: centroids descending
            cents = np.sort(np.exp(rng.uniform(np.log(0.01), np.log(0.4), p // 2)))[::-1]
            widths = np.exp(rng.uniform(np.log(0.003), np.log(0.2), (p + 1) // 2))
            for i in range(p // 2):
                w = complex(-2 * np.pi * widths[i], 2 * np.pi * cents[i])
                th += [np.log(abs(w) ** 2), np.log(-2 * w.real)]
            if p % 2:
                th.append(np.log(2 * np.pi * widths[-1]))
            th += list(np.abs(rng.standard_normal(q)))
            rows.append(th)
        # one row with an overdamped (real) AR pair
        rows[-1][3] = np.log(0.02)
        rows[-1][4] = np.log(0.9)
        rows = np.array(rows)
        lp = np.array([ref_logpost(th, p, q, ts, ys, es) for th in rows])
        cases["c%d%d" % (p, q)] = dict(p=p, q=q, theta=rows, logpost=lp[:, 0], loglik=lp[:, 1], logprior=lp[:, 2])
        print("c%d%d logpost[0] = %.10f" % (p, q, lp[0, 0]))
    # ---- ZCARMA(5): kappa bounds as carpack.hpp:413-419
    dts = np.diff(ts)
    khi = 1.0 / dts.min()
    klo = max(1.0 / (ts.max() - ts.min()), 1.0 / (10.0 * np.median(dts)))
    rows = []
    for k in range(6):
        th = list(base.copy())
        th[0] *= np.exp(0.1 * rng.standard_normal())
        th[3:] = list(np.array(th[3:]) + 0.2 * rng.standard_normal(5))
        th.append(rng.uniform(-2, 2))
        rows.append(th)
    rows = np.array(rows)
    lp = np.array([ref_logpost(th, 5, 0, ts, ys, es, zcarma_kappa_bounds=(klo, khi)) for th in rows])
    cases["z5"] = dict(p=5, q=0, theta=rows, logpost=lp[:, 0], loglik=lp[:, 1], logprior=lp[:, 2],
                       kappa_bounds=np.array([klo, khi]))
    flat = dict(t270=t, y270=y, ysig270=ysig, t60=ts, y60=ys, ysig60=es, true_roots=ar_roots,
                true_ma=ma_true, true_sigsqr=sigsqr)
    for k, v in cases.items():
        for kk, vv in v.items():
            flat["%s_%s" % (k, kk)] = np.asarray(vv)
    np.savez(os.path.join(HERE, "loglik_cases.npz"), **flat)


def make_derived_fixture():
    """Posterior post-processing of CarmaSample (src/carmcmc/carma_pack.py:439-546: _ar_roots, _ma_coefs, _sigma_noise,
    run unchanged) on the theta rows of loglik_cases.npz: AR roots, MA coefficients and sigma^2 per row."""
    cases = dict(np.load(os.path.join(HERE, "loglik_cases.npz")))
    out = {}
    for name in sorted({k.split("_")[0] for k in cases if k.startswith("c") and k.endswith("_theta")}):
        p, q = int(cases[name + "_p"]), int(cases[name + "_q"])
        roots, ma, sigsqr = theta_to_params(cases[name + "_theta"], p, q)
        out[name + "_roots"], out[name + "_ma"], out[name + "_sigsqr"] = roots, ma, np.ravel(sigsqr)
    np.savez_compressed(os.path.join(HERE, "derived_params.npz"), **out)


def make_mcmc_fixtures():
    """The reference's own 1000-point test series for the long statistical tests (posterior recovery within 3 sigma,
    residual whiteness): cpp_tests/data/{car5,zcar5,carma}_test.dat with the true parameters of
    cpp_tests/carma_unit_tests.cpp:1378-1656 / generate_test_data.py:47-58.  Data only (time, y, yerr)."""
    out = {}
    for name in ("car5", "zcar5", "carma"):
        data = np.loadtxt(os.path.join(REF, "cpp_tests/data/%s_test.dat" % name))
        out[name + "_t"], out[name + "_y"], out[name + "_yerr"] = data[:, 0], data[:, 1], data[:, 2]
    out["qpo_width"] = np.array([0.01, 0.01, 0.002])
    out["qpo_cent"] = np.array([0.2, 0.02])
    out["sigmay"], out["kappa"] = np.array(2.3), np.array(0.5)
    np.savez_compressed(os.path.join(HERE, "mcmc_fixtures.npz"), **out)


def make_car1_fixture():
    """CAR(1): the reference has no numpy CAR(1) filter (KalmanFilter1 is C++-only and
    KalmanFilterDeprecated indexes row 1 of a 1x1 matrix), so CAR(1) is pinned the way the
    reference's own tests pin it (carma_unit_tests.cpp:277-384, brute-force Gaussian process):
    the OU covariance sigma_y^2 exp(-omega |dt|) + diag(scale yerr^2) is factorised densely and the
    exact log-likelihood  -1/2 log det K - 1/2 r^T K^-1 r  (no 2 pi term, carpack.hpp:167-171)
    is stored, together with the one-step predictive means/variances from the Cholesky factor."""
    data = np.loadtxt(os.path.join(REF, "cpp_tests/data/car1_test.dat"))
    t, y, yerr = data[:200, 0].copy(), data[:200, 1].copy(), data[:200, 2].copy()
    rows, lps, means, variances = [], [], [], []
    rng = np.random.default_rng(7)
    for k in range(8):
        sig_y = np.std(y) * np.exp(0.2 * rng.standard_normal())
        scale = np.clip(1 + 0.2 * rng.standard_normal(), 0.55, 1.9)
        mu = np.mean(y) + 0.1 * rng.standard_normal()
        logw = np.log(1.0 / (np.median(np.diff(t)) * rng.uniform(1, 50)))
        omega = np.exp(logw)
        K = sig_y ** 2 * np.exp(-omega * np.abs(t[:, None] - t[None, :])) + np.diag(scale * yerr ** 2)
        L = np.linalg.cholesky(K)
        r = y - mu
        e = np.linalg.solve(L, r)            # standardised innovations
        var = np.diag(L) ** 2                 # one-step predictive variances
        mean = r - e * np.diag(L)             # one-step predictive means
        ll = np.sum(-0.5 * np.log(var) - 0.5 * e ** 2)
        prior = -0.5 * 50.0 / scale - 26.0 * np.log(scale)
        rows.append([sig_y, scale, mu, logw])
        lps.append(ll + prior)
        means.append(mean)
        variances.append(var)
    np.savez(os.path.join(HERE, "car1_cases.npz"), t=t, y=y, yerr=yerr, theta=np.array(rows), logpost=np.array(lps),
             mean=np.array(means), var=np.array(variances))
    print("car1 logpost[0] = %.10f" % lps[0])


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (build container only)")
    make_kelly_fixture()
    make_loglik_fixture()
    make_car1_fixture()
    make_mcmc_fixtures()
    make_derived_fixture()
