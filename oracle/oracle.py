"""ctypes front-end of the CPU oracle (oracle/carma_oracle.cpp).

TEST INFRASTRUCTURE ONLY: may be imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product (carma_pack_b200) never
imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

KIND_CAR1, KIND_CARP, KIND_CARMA, KIND_ZCAR, KIND_ZCARMA = 0, 1, 2, 3, 4


class Prior(ctypes.Structure):
    _fields_ = [("max_stdev", ctypes.c_double), ("max_freq", ctypes.c_double), ("min_freq", ctypes.c_double),
                ("kappa_low", ctypes.c_double), ("kappa_high", ctypes.c_double), ("measerr_dof", ctypes.c_double)]


class PTOpts(ctypes.Structure):
    _fields_ = [("nsamples", ctypes.c_int), ("burnin", ctypes.c_int), ("thin", ctypes.c_int), ("ntemps", ctypes.c_int),
                ("tmax", ctypes.c_double), ("dof", ctypes.c_int), ("target_rate", ctypes.c_double),
                ("gamma", ctypes.c_double), ("seed", ctypes.c_uint64), ("ensemble", ctypes.c_uint32),
                ("max_start_attempts", ctypes.c_int)]


TRACE_DTYPE = np.dtype([("lp_prop", "f8"), ("lp_cur", "f8"), ("alpha", "f8"), ("u", "f8"),
                        ("accepted", "i4"), ("pad", "i4")])

_dp = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    """Compile liboracle.so / liboracle_fast.so with the committed Makefile."""
    need = force or not all(os.path.exists(os.path.join(_HERE, n)) for n in ("liboracle.so", "liboracle_fast.so"))
    if need:
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)


_libs = {}


def lib(fast=False):
    name = "liboracle_fast.so" if fast else "liboracle.so"
    if name not in _libs:
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.oracle_variance.restype = ctypes.c_double
        L.oracle_log_prior.restype = ctypes.c_double
        L.oracle_tdist.restype = ctypes.c_double
        L.oracle_tdist.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int]
        L.oracle_philox.argtypes = [ctypes.c_uint32] * 4 + [ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint32)]
        _libs[name] = L
    return _libs[name]


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def model_dim(kind, p, q):
    if kind == KIND_CAR1:
        return 4
    if kind == KIND_CARMA:
        return 3 + p + q
    if kind == KIND_ZCARMA:
        return 4 + p
    return 3 + p


def default_prior(t, y, population_var=True):
    t_, tp = _d(t)
    y_, yp = _d(y)
    pr = Prior()
    lib().oracle_default_prior(tp, yp, ctypes.c_size_t(t_.size), int(population_var), ctypes.byref(pr))
    return pr


def ar_roots(logq):
    q_, qp = _d(logq)
    out = np.empty(2 * q_.size)
    lib().oracle_ar_roots(qp, int(q_.size), out.ctypes.data_as(_dp))
    return out[0::2] + 1j * out[1::2]


def ma_coefs(kind, theta, p, q, prior):
    th, thp = _d(theta)
    out = np.empty(p)
    lib().oracle_ma_coefs(kind, thp, p, q, ctypes.byref(prior), out.ctypes.data_as(_dp))
    return out


def variance(roots, ma, sigma=1.0, lag=0.0):
    roots = np.asarray(roots, dtype=complex)
    r = np.empty(2 * roots.size)
    r[0::2] = roots.real
    r[1::2] = roots.imag
    m, mp = _d(ma)
    return lib().oracle_variance(r.ctypes.data_as(_dp), mp, int(roots.size), int(m.size),
                                 ctypes.c_double(sigma), ctypes.c_double(lag))


def check_prior(kind, theta, p, prior, ignore_prior=False):
    th, thp = _d(theta)
    return bool(lib().oracle_check_prior(kind, thp, p, ctypes.byref(prior), int(ignore_prior)))


def log_prior(kind, theta, p, prior):
    th, thp = _d(theta)
    return lib().oracle_log_prior(kind, thp, p, ctypes.byref(prior))


def _omega_buf(omega):
    omega = np.asarray(omega, dtype=complex)
    r = np.empty(2 * omega.size)
    r[0::2] = omega.real
    r[1::2] = omega.imag
    return r


def filterp(t, y, yerr, sigsqr, omega, ma):
    """KalmanFilterp(t, y, yerr, sigsqr, omega, ma).Filter(); returns (mean, var)."""
    t_, tp = _d(t)
    y_, yp = _d(y)
    e_, ep = _d(yerr)
    om = _omega_buf(omega)
    p = om.size // 2
    m = np.zeros(p)
    m[:len(ma)] = ma
    mean = np.empty(t_.size)
    var = np.empty(t_.size)
    rc = lib().oracle_filterp(tp, yp, ep, ctypes.c_size_t(t_.size), ctypes.c_double(sigsqr),
                              om.ctypes.data_as(_dp), m.ctypes.data_as(_dp), p,
                              mean.ctypes.data_as(_dp), var.ctypes.data_as(_dp))
    if rc:
        raise RuntimeError("singular Vandermonde solve")
    return mean, var


def predictp(t, y, yerr, sigsqr, omega, ma, tq):
    t_, tp = _d(t)
    y_, yp = _d(y)
    e_, ep = _d(yerr)
    q_, qp = _d(np.atleast_1d(tq))
    om = _omega_buf(omega)
    p = om.size // 2
    m = np.zeros(p)
    m[:len(ma)] = ma
    qm = np.empty(q_.size)
    qv = np.empty(q_.size)
    rc = lib().oracle_predictp(tp, yp, ep, ctypes.c_size_t(t_.size), ctypes.c_double(sigsqr),
                               om.ctypes.data_as(_dp), m.ctypes.data_as(_dp), p, qp, ctypes.c_size_t(q_.size),
                               qm.ctypes.data_as(_dp), qv.ctypes.data_as(_dp))
    if rc:
        raise RuntimeError("singular Vandermonde solve")
    return qm, qv


def filter1(t, y, yerr, sigsqr, omega):
    t_, tp = _d(t)
    y_, yp = _d(y)
    e_, ep = _d(yerr)
    mean = np.empty(t_.size)
    var = np.empty(t_.size)
    lib().oracle_filter1(tp, yp, ep, ctypes.c_size_t(t_.size), ctypes.c_double(sigsqr), ctypes.c_double(omega),
                         mean.ctypes.data_as(_dp), var.ctypes.data_as(_dp))
    return mean, var


def predict1(t, y, yerr, sigsqr, omega, tq):
    t_, tp = _d(t)
    y_, yp = _d(y)
    e_, ep = _d(yerr)
    q_, qp = _d(np.atleast_1d(tq))
    qm = np.empty(q_.size)
    qv = np.empty(q_.size)
    lib().oracle_predict1(tp, yp, ep, ctypes.c_size_t(t_.size), ctypes.c_double(sigsqr), ctypes.c_double(omega),
                          qp, ctypes.c_size_t(q_.size), qm.ctypes.data_as(_dp), qv.ctypes.data_as(_dp))
    return qm, qv


def logdensity(kind, p, q, t, y, yerr, theta, prior=None, ignore_prior=False, long_double=False, fast=False, lean=False):
    """CARMA_Base::LogDensity for each row of theta (n x d).  lean: the fixed-size, heap-free variant (p >= 2)."""
    t_, tp = _d(t)
    y_, yp = _d(y)
    e_, ep = _d(yerr)
    th = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
    d = model_dim(kind, p, q)
    assert th.shape[1] == d, (th.shape, d)
    if prior is None:
        prior = default_prior(t_, y_)
    out = np.empty(th.shape[0])
    L = lib(fast)
    fn = L.oracle_logdensity_batch_ld if long_double else (L.oracle_logdensity_batch_lean if lean else L.oracle_logdensity_batch)
    fn(kind, p, q, tp, yp, ep, ctypes.c_size_t(t_.size), ctypes.byref(prior), int(ignore_prior),
       th.ctypes.data_as(_dp), ctypes.c_size_t(th.shape[0]), out.ctypes.data_as(_dp))
    return out


def logdensity_multi(kind, p, q, t, y, yerr, offsets, theta, priors, ignore_prior=False, fast=False):
    t_, tp = _d(t)
    y_, yp = _d(y)
    e_, ep = _d(yerr)
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    nc = off.size - 1
    th = np.ascontiguousarray(theta, dtype=np.float64)
    arr = (Prior * nc)(*priors)
    out = np.empty(nc)
    lib(fast).oracle_logdensity_multi(kind, p, q, tp, yp, ep, off.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                      ctypes.c_size_t(nc), arr, int(ignore_prior), th.ctypes.data_as(_dp),
                                      out.ctypes.data_as(_dp))
    return out


def philox(c0, c1, c2, c3, seed):
    out = (ctypes.c_uint32 * 4)()
    lib().oracle_philox(c0, c1, c2, c3, ctypes.c_uint64(seed), out)
    return [int(x) for x in out]


def tdist(seed, chain, it, j, dof=8):
    return lib().oracle_tdist(ctypes.c_uint64(seed), chain, it, j, dof)


def chol_update(L, v, downdate):
    L = np.array(L, dtype=np.float64, order="C")
    v = np.array(v, dtype=np.float64)
    lib().oracle_chol_update(L.ctypes.data_as(_dp), v.ctypes.data_as(_dp), int(L.shape[0]), int(downdate))
    return L, v


def starting_value(kind, p, q, t, y, yerr, prior, seed, chain, max_attempts=1000):
    t_, tp = _d(t)
    y_, yp = _d(y)
    e_, ep = _d(yerr)
    d = model_dim(kind, p, q)
    th = np.empty(d)
    lp = ctypes.c_double()
    att = lib().oracle_starting_value(kind, p, q, tp, yp, ep, ctypes.c_size_t(t_.size), ctypes.byref(prior),
                                      ctypes.c_uint64(seed), ctypes.c_uint32(chain), max_attempts,
                                      th.ctypes.data_as(_dp), ctypes.byref(lp))
    return th, lp.value, att


def pt_run(kind, p, q, t, y, yerr, nsamples, burnin, thin=1, ntemps=10, seed=1, ensemble=0, init=None, prior=None,
           tmax=100.0, dof=8, target_rate=0.25, gamma=2.0 / 3.0, max_start_attempts=1000, want_trace=False,
           fast=False):
    """RunCarmaSampler restated (carmcmc.cpp:79-177) with Philox streams."""
    t_, tp = _d(t)
    y_, yp = _d(y)
    e_, ep = _d(yerr)
    if prior is None:
        prior = default_prior(t_, y_)
    d = model_dim(kind, p, q)
    o = PTOpts(nsamples, burnin, thin, ntemps, tmax, dof, target_rate, gamma, seed, ensemble, max_start_attempts)
    samples = np.empty((nsamples, d))
    logposts = np.empty(nsamples)
    rates = np.empty(ntemps)
    iters = burnin + nsamples * thin
    trace = np.zeros((2, iters, ntemps), dtype=TRACE_DTYPE) if want_trace else None
    ttrace = np.zeros((iters, ntemps, d)) if want_trace else None
    chol = np.empty((ntemps, d, d))
    initp = None
    if init is not None and len(init) == d:
        init_, initp = _d(init)
    rc = lib(fast).oracle_pt_run(kind, p, q, tp, yp, ep, ctypes.c_size_t(t_.size), ctypes.byref(prior),
                                 ctypes.byref(o), initp, samples.ctypes.data_as(_dp), logposts.ctypes.data_as(_dp),
                                 rates.ctypes.data_as(_dp),
                                 trace.ctypes.data_as(ctypes.c_void_p) if want_trace else None,
                                 ttrace.ctypes.data_as(_dp) if want_trace else None,
                                 chol.ctypes.data_as(_dp))
    if rc:
        raise RuntimeError("oracle_pt_run failed rc=%d" % rc)
    res = dict(samples=samples, logposts=logposts, accept_rates=rates, chol=chol)
    if want_trace:
        res["ram_trace"] = trace[0]
        res["exchange_trace"] = trace[1]
        res["proposals"] = ttrace
    return res
