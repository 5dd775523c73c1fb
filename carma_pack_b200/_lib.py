"""ctypes binding of the C ABI in include/carma_b200.h (libcarma_b200.so).

This is the only way Python reaches the GPU path.  There is no CPU fallback: if the shared
library is missing the import fails loudly, and every compute call raises CarmaError when CUDA is
unavailable.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CARMA_B200_LIB") or os.path.join(_HERE, "libcarma_b200.so")  # override: tuning builds


def _point_at_bundled_nccl():
    """The library loads NCCL lazily (carma_comm_* / carma_gather_summaries).  When this Python environment ships a
    pip-bundled NCCL (the one torch links against), make that the copy that gets mapped, so that a later
    `import torch` in the same process finds the symbols it needs behind the soname libnccl.so.2.  torch itself is not
    imported here."""
    if os.environ.get("CARMA_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for loc in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(loc, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["CARMA_NCCL_LIB"] = cand
                return
    except Exception:  # noqa: BLE001 - best effort: the system NCCL is the fallback
        pass


_point_at_bundled_nccl()

KIND_CAR1, KIND_CARP, KIND_CARMA, KIND_ZCAR, KIND_ZCARMA = 0, 1, 2, 3, 4
IGNORE_BOUNDS, LOGLIK_ONLY = 1, 2
MAX_P = 7

EXPORTED_SYMBOLS = [
    "carma_last_error", "carma_abi_version", "carma_device_count",
    "carma_series_create", "carma_series_destroy", "carma_series_length", "carma_series_default_prior",
    "carma_loglik_batch_dev", "carma_loglik_batch", "carma_loglik_batch_async", "carma_loglik_batch_wait",
    "carma_log_prior", "carma_loglik_scan_dev", "carma_loglik_scan",
    "carma_multi_series_create", "carma_multi_series_destroy", "carma_multi_series_default_priors",
    "carma_multi_series_simulate", "carma_multi_series_get_curve",
    "carma_mle_default_opts", "carma_mle_batch", "carma_lbfgs_batch",
    "carma_multi_loglik_dev", "carma_multi_loglik",
    "carma_filter", "carma_predict",
    "carma_pt_default_opts", "carma_pt_run", "carma_pt_run_dev", "carma_multi_pt_run",
    "carma_fp64_peak_tflops", "carma_philox_dev", "carma_tdist_dev", "carma_fastmath_dev",
    "carma_simulate", "carma_starting_value", "carma_comm_unique_id", "carma_comm_init_rank", "carma_comm_destroy",
    "carma_gather_summaries", "carma_gather_summaries_dev", "carma_derived_params", "carma_derived_params_dev", "carma_mle_batch_device",
    "carma_mle_grid_device",
]


class CarmaError(RuntimeError):
    pass


class Prior(ctypes.Structure):
    _fields_ = [("max_stdev", ctypes.c_double), ("max_freq", ctypes.c_double), ("min_freq", ctypes.c_double),
                ("kappa_low", ctypes.c_double), ("kappa_high", ctypes.c_double), ("measerr_dof", ctypes.c_double)]

    def as_tuple(self):
        return (self.max_stdev, self.max_freq, self.min_freq, self.kappa_low, self.kappa_high, self.measerr_dof)


PRIOR_DTYPE = np.dtype([("max_stdev", "f8"), ("max_freq", "f8"), ("min_freq", "f8"), ("kappa_low", "f8"),
                        ("kappa_high", "f8"), ("measerr_dof", "f8")])


class PTOpts(ctypes.Structure):
    _fields_ = [("nsamples", ctypes.c_int), ("burnin", ctypes.c_int), ("thin", ctypes.c_int),
                ("ntemps", ctypes.c_int), ("tmax", ctypes.c_double), ("dof", ctypes.c_int),
                ("target_rate", ctypes.c_double), ("gamma", ctypes.c_double), ("seed", ctypes.c_uint64),
                ("ensemble_offset", ctypes.c_uint32), ("max_start_attempts", ctypes.c_int),
                ("order_mode", ctypes.c_int), ("record_trace", ctypes.c_int)]


class MLEJob(ctypes.Structure):   # carma_mle_job_t
    _fields_ = [("kind", ctypes.c_int), ("p", ctypes.c_int), ("q", ctypes.c_int), ("flags", ctypes.c_uint),
                ("prior", Prior), ("nstart", ctypes.c_size_t)]


MAX_DIM = 17   # CARMA_MAX_DIM


class MLEOpts(ctypes.Structure):
    _fields_ = [("maxiter", ctypes.c_int), ("history", ctypes.c_int), ("max_backtrack", ctypes.c_int),
                ("reserved", ctypes.c_int), ("gtol", ctypes.c_double), ("ftol", ctypes.c_double),
                ("fd_eps", ctypes.c_double)]


TRACE_DTYPE = np.dtype([("lp_prop", "f8"), ("lp_cur", "f8"), ("alpha", "f8"), ("u", "f8"),
                        ("accepted", "i4"), ("pad", "i4")])

_dp = ctypes.POINTER(ctypes.c_double)
_vp = ctypes.c_void_p
_sz = ctypes.c_size_t
OBJECTIVE_FN = ctypes.CFUNCTYPE(ctypes.c_int, _dp, _sz, _sz, _dp, _vp)


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "carma_pack_b200: %s is missing. Build it with `python build_native.py` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    L.carma_last_error.restype = ctypes.c_char_p
    pr = ctypes.POINTER(Prior)
    L.carma_series_create.argtypes = [_dp, _dp, _dp, _sz, ctypes.c_int, ctypes.POINTER(_vp)]
    L.carma_series_destroy.argtypes = [_vp]
    L.carma_series_length.argtypes = [_vp, ctypes.POINTER(_sz)]
    L.carma_series_default_prior.argtypes = [_vp, ctypes.c_int, pr]
    L.carma_loglik_batch_dev.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, pr, _sz, _vp, _vp,
                                         ctypes.c_uint, _vp]
    L.carma_loglik_batch.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, pr, _sz, _vp, _vp, ctypes.c_uint]
    L.carma_loglik_scan_dev.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, pr, _sz, _vp, _vp,
                                        ctypes.c_uint, ctypes.c_int, _vp]
    L.carma_loglik_scan.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, pr, _sz, _vp, _vp, ctypes.c_uint,
                                    ctypes.c_int]
    L.carma_loglik_batch_async.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, pr, _sz, _vp, _vp,
                                           ctypes.c_uint, ctypes.c_int]
    L.carma_loglik_batch_wait.argtypes = [_vp, ctypes.c_int]
    L.carma_log_prior.argtypes = [ctypes.c_int, ctypes.c_int, _dp, pr, _dp]
    L.carma_multi_series_create.argtypes = [_dp, _dp, _dp, ctypes.POINTER(ctypes.c_int64), _sz, ctypes.c_int,
                                            ctypes.POINTER(_vp)]
    L.carma_multi_series_destroy.argtypes = [_vp]
    L.carma_multi_series_default_priors.argtypes = [_vp, ctypes.c_int, _vp]
    L.carma_multi_series_simulate.argtypes = [_sz, _sz, ctypes.c_int, ctypes.c_int, ctypes.c_int, _dp, pr,
                                              ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_uint64,
                                              ctypes.c_uint32, ctypes.c_int, ctypes.POINTER(_vp)]
    L.carma_multi_series_get_curve.argtypes = [_vp, _sz, _dp, _dp, _dp, _sz, ctypes.POINTER(_sz)]
    L.carma_mle_default_opts.argtypes = [ctypes.POINTER(MLEOpts)]
    L.carma_mle_default_opts.restype = None
    L.carma_lbfgs_batch.argtypes = [OBJECTIVE_FN, _vp, _sz, _sz, _dp, _dp, _dp, ctypes.POINTER(MLEOpts), _dp, _dp,
                                    ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_longlong)]
    L.carma_mle_batch.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, pr, ctypes.c_uint, _sz, _dp, _dp, _dp,
                                  ctypes.POINTER(MLEOpts), _dp, _dp, ctypes.POINTER(ctypes.c_int),
                                  ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
    L.carma_mle_batch_device.argtypes = L.carma_mle_batch.argtypes
    L.carma_mle_grid_device.argtypes = [_vp, ctypes.c_int, ctypes.POINTER(MLEJob), _dp, _dp, _dp, ctypes.POINTER(MLEOpts),
                                        _dp, _dp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
    L.carma_multi_loglik_dev.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp,
                                         ctypes.c_uint, _vp]
    L.carma_multi_loglik.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, ctypes.c_uint]
    L.carma_filter.argtypes = [_vp, ctypes.c_double, _dp, _dp, ctypes.c_int, ctypes.c_double, ctypes.c_double, _dp, _dp]
    L.carma_predict.argtypes = [_vp, ctypes.c_double, _dp, _dp, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                _dp, _sz, _dp, _dp]
    L.carma_pt_default_opts.argtypes = [ctypes.POINTER(PTOpts)]
    L.carma_pt_default_opts.restype = None
    L.carma_pt_run.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, pr, ctypes.POINTER(PTOpts), _sz,
                               _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    L.carma_pt_run_dev.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, pr, ctypes.POINTER(PTOpts), _sz,
                                   _vp, _vp, _vp, _vp, _vp, _vp]
    L.carma_multi_pt_run.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, ctypes.POINTER(PTOpts), _sz,
                                     _vp, _vp, _vp, _vp]
    L.carma_fp64_peak_tflops.argtypes = [ctypes.c_int, _dp]
    L.carma_philox_dev.argtypes = [ctypes.c_uint32] * 4 + [ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint32)]
    L.carma_fastmath_dev.argtypes = [_dp, _dp, _sz, _dp, _dp, _dp, _dp, _dp, _dp]
    L.carma_simulate.argtypes = [_vp, ctypes.c_double, _dp, _dp, ctypes.c_int, ctypes.c_double, ctypes.c_double, _dp, _sz,
                                 ctypes.c_uint64, _sz, _dp]
    L.carma_starting_value.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(Prior), ctypes.c_uint64,
                                       ctypes.c_uint32, ctypes.c_int, _dp, _dp]
    L.carma_derived_params.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, _sz, _dp, _dp, ctypes.c_int]
    L.carma_derived_params_dev.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, _sz, _vp, _vp, _vp]
    L.carma_comm_unique_id.argtypes = [ctypes.c_char_p]
    L.carma_comm_init_rank.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    L.carma_comm_destroy.argtypes = [_vp]
    L.carma_gather_summaries.argtypes = [_vp, _dp, _sz, _dp, _vp]
    L.carma_gather_summaries_dev.argtypes = [_vp, _vp, _sz, _vp, _vp]
    L.carma_tdist_dev.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, _dp]
    return L


lib = _load()


def check(rc, what=""):
    if rc != 0:
        msg = lib.carma_last_error()
        raise CarmaError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else ""))


def model_dim(kind, p, q):
    if kind == KIND_CAR1:
        return 4
    if kind == KIND_CARMA:
        return 3 + p + q
    if kind == KIND_ZCARMA:
        return 4 + p
    return 3 + p


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return a.ctypes.data_as(_dp)


def device_count():
    n = ctypes.c_int(0)
    rc = lib.carma_device_count(ctypes.byref(n))
    return n.value if rc == 0 else 0


def derived_params(kind, p, q, theta, prior=None, device=0):
    """Posterior post-processing on the device (carma_derived_params): theta rows -> dict of ar_roots (complex, n x p),
    ar_coefs (n x p+1, highest power first), ma_coefs (n x p, beta_0 = 1, zero beyond q), sigma (n), psd_width and
    psd_centroid (n x p)."""
    th = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
    d = model_dim(kind, p, q)
    if th.shape[1] != d:
        raise ValueError("theta must have %d columns for this model, got %d" % (d, th.shape[1]))
    n = th.shape[0]
    out = np.empty((n, 6 * p + 2))
    check(lib.carma_derived_params(kind, p, q, ctypes.byref(prior) if prior is not None else None, n, _ptr(th), _ptr(out),
                                   device), "carma_derived_params")
    return {"ar_roots": out[:, 0:2 * p:2] + 1j * out[:, 1:2 * p:2], "ar_coefs": out[:, 2 * p:3 * p + 1].copy(),
            "ma_coefs": out[:, 3 * p + 1:4 * p + 1].copy(), "sigma": out[:, 4 * p + 1].copy(),
            "psd_width": out[:, 4 * p + 2:5 * p + 2].copy(), "psd_centroid": out[:, 5 * p + 2:6 * p + 2].copy()}


class Series:
    """One light curve resident in HBM (carma_series_t)."""

    def __init__(self, time, y, yerr, device=0):
        t, yy, ee = _c(time), _c(y), _c(yerr)
        if not (t.shape == yy.shape == ee.shape and t.ndim == 1):
            raise ValueError("time, y, yerr must be 1-d arrays of equal length")
        self.handle = _vp()
        check(lib.carma_series_create(_ptr(t), _ptr(yy), _ptr(ee), t.size, device, ctypes.byref(self.handle)),
              "carma_series_create")
        self.ny = t.size
        self.device = device
        self.time, self.y, self.yerr = t, yy, ee

    def close(self):
        if getattr(self, "handle", None):
            lib.carma_series_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def default_prior(self, population_var=True):
        pr = Prior()
        check(lib.carma_series_default_prior(self.handle, int(population_var), ctypes.byref(pr)), "default_prior")
        return pr

    def loglik(self, kind, p, q, theta, prior=None, flags=0):
        """Batched CARMA_Base::LogDensity with host arrays (H2D + kernel + D2H inside)."""
        th = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
        d = model_dim(kind, p, q)
        if th.shape[1] != d:
            raise ValueError("theta must have %d columns for this model, got %d" % (d, th.shape[1]))
        if prior is None:
            prior = self.default_prior()
        out = np.empty(th.shape[0])
        check(lib.carma_loglik_batch(self.handle, kind, p, q, ctypes.byref(prior), th.shape[0],
                                     th.ctypes.data, out.ctypes.data, flags), "carma_loglik_batch")
        return out

    def loglik_async(self, kind, p, q, theta_ptr, out_ptr, n, prior, slot, flags=0):
        """Enqueue H2D + kernel + D2H for host buffers (raw addresses, ideally pinned) on pipeline slot 0/1."""
        check(lib.carma_loglik_batch_async(self.handle, kind, p, q, ctypes.byref(prior), n, theta_ptr, out_ptr, flags,
                                           slot), "carma_loglik_batch_async")

    def loglik_wait(self, slot):
        check(lib.carma_loglik_batch_wait(self.handle, slot), "carma_loglik_batch_wait")

    def loglik_scan(self, kind, p, q, theta, prior=None, flags=0, chunk=0):
        """Same value as loglik(), computed by the temporally parallel scan (for one very long series)."""
        th = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
        d = model_dim(kind, p, q)
        if th.shape[1] != d:
            raise ValueError("theta must have %d columns for this model, got %d" % (d, th.shape[1]))
        if prior is None:
            prior = self.default_prior()
        out = np.empty(th.shape[0])
        check(lib.carma_loglik_scan(self.handle, kind, p, q, ctypes.byref(prior), th.shape[0], th.ctypes.data,
                                    out.ctypes.data, flags, chunk), "carma_loglik_scan")
        return out

    def loglik_scan_dev(self, kind, p, q, d_theta_ptr, d_out_ptr, n, prior, flags=0, chunk=0, stream=0):
        check(lib.carma_loglik_scan_dev(self.handle, kind, p, q, ctypes.byref(prior), n, d_theta_ptr, d_out_ptr, flags,
                                        chunk, stream), "carma_loglik_scan_dev")

    def loglik_dev(self, kind, p, q, d_theta_ptr, d_out_ptr, n, prior, flags=0, stream=0):
        """Device-resident variant: raw device pointers (e.g. torch tensor .data_ptr()), no sync."""
        check(lib.carma_loglik_batch_dev(self.handle, kind, p, q, ctypes.byref(prior), n, d_theta_ptr, d_out_ptr,
                                         flags, stream), "carma_loglik_batch_dev")

    def filter(self, sigsqr, omega, ma, measerr_scale=1.0, mu=0.0):
        om = np.asarray(omega, dtype=complex).ravel()
        p = om.size
        buf = np.empty(2 * p)
        buf[0::2], buf[1::2] = om.real, om.imag
        m = np.zeros(p)
        m[:len(ma)] = ma
        mean, var = np.empty(self.ny), np.empty(self.ny)
        check(lib.carma_filter(self.handle, sigsqr, _ptr(buf), _ptr(m), p, measerr_scale, mu, _ptr(mean), _ptr(var)),
              "carma_filter")
        return mean, var

    def predict(self, sigsqr, omega, ma, tq, measerr_scale=1.0, mu=0.0):
        om = np.asarray(omega, dtype=complex).ravel()
        p = om.size
        buf = np.empty(2 * p)
        buf[0::2], buf[1::2] = om.real, om.imag
        m = np.zeros(p)
        m[:len(ma)] = ma
        q = _c(np.atleast_1d(tq))
        qm, qv = np.empty(q.size), np.empty(q.size)
        check(lib.carma_predict(self.handle, sigsqr, _ptr(buf), _ptr(m), p, measerr_scale, mu, _ptr(q), q.size,
                                _ptr(qm), _ptr(qv)), "carma_predict")
        return qm, qv

    def simulate(self, sigsqr, omega, ma, tsim, measerr_scale=1.0, mu=0.0, seed=1, npaths=1):
        """Conditional simulation of the process at `tsim` given the data (carma_simulate): (npaths, nsim) array."""
        om = np.asarray(omega, dtype=complex).ravel()
        p = om.size
        buf = np.empty(2 * p)
        buf[0::2], buf[1::2] = om.real, om.imag
        m = np.zeros(p)
        m[:len(ma)] = ma
        q = _c(np.atleast_1d(tsim))
        out = np.empty((int(npaths), q.size))
        check(lib.carma_simulate(self.handle, sigsqr, _ptr(buf), _ptr(m), p, measerr_scale, mu, _ptr(q), q.size,
                                 int(seed), int(npaths), _ptr(out)), "carma_simulate")
        return out

    def pt_run(self, kind, p, q, nsamples, burnin, thin=1, ntemps=10, n_ensembles=1, seed=1, ensemble_offset=0,
               init=None, prior=None, order_mode=0, record_trace=False, tmax=100.0, dof=8, target_rate=0.25,
               gamma=2.0 / 3.0, max_start_attempts=1000):
        """RunCarmaSampler for n_ensembles independent ensembles, fully on device."""
        d = model_dim(kind, p, q)
        if prior is None:
            prior = self.default_prior()
        o = PTOpts()
        lib.carma_pt_default_opts(ctypes.byref(o))
        o.nsamples, o.burnin, o.thin, o.ntemps = int(nsamples), int(burnin), int(thin), int(ntemps)
        o.tmax, o.dof, o.target_rate, o.gamma = tmax, dof, target_rate, gamma
        o.seed, o.ensemble_offset, o.max_start_attempts = seed, ensemble_offset, max_start_attempts
        o.order_mode, o.record_trace = order_mode, int(record_trace)
        samples = np.empty((n_ensembles, nsamples, d))
        logposts = np.empty((n_ensembles, nsamples))
        acc = np.empty((n_ensembles, ntemps))
        xr = np.empty((n_ensembles, ntemps))
        iters = burnin + nsamples * thin
        rt = xt = prop = None
        if record_trace:
            rt = np.zeros((n_ensembles, iters, ntemps), dtype=TRACE_DTYPE)
            xt = np.zeros((n_ensembles, iters, ntemps), dtype=TRACE_DTYPE)
            prop = np.zeros((n_ensembles, iters, ntemps, d))
        initp = None
        if init is not None and len(init) == d:
            init = _c(init)
            initp = init.ctypes.data
        check(lib.carma_pt_run(self.handle, kind, p, q, ctypes.byref(prior), ctypes.byref(o), n_ensembles, initp,
                               samples.ctypes.data, logposts.ctypes.data, acc.ctypes.data, xr.ctypes.data,
                               rt.ctypes.data if record_trace else None, xt.ctypes.data if record_trace else None,
                               prop.ctypes.data if record_trace else None), "carma_pt_run")
        res = dict(samples=samples, logposts=logposts, accept_rates=acc, exchange_rates=xr)
        if record_trace:
            res.update(ram_trace=rt, exchange_trace=xt, proposals=prop)
        return res

    def mle_batch(self, kind, p, q, x0, lower, upper, prior=None, flags=0, maxiter=1000, history=8, gtol=1e-5,
                  ftol=2.2e-9, fd_eps=1e-8, slot=0, on_device=False):
        """Projected L-BFGS from every row of x0: minimises -LogDensity over the box [lower, upper].  on_device=False:
        carma_mle_batch (host loop, all rows in lock-step, one launch per batch of trial points); on_device=True:
        carma_mle_batch_device (the whole fit of a start inside one kernel, one warp per start).  Returns
        (x, f, nit, nfev).  Releases the GIL for the whole fit."""
        d = model_dim(kind, p, q)
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        if x0.ndim != 2 or x0.shape[1] != d:
            raise ValueError("x0 must be (nstart, %d)" % d)
        lo, hi = _c(np.broadcast_to(lower, (d,))), _c(np.broadcast_to(upper, (d,)))
        if prior is None:
            prior = self.default_prior()
        o = MLEOpts()
        lib.carma_mle_default_opts(ctypes.byref(o))
        o.maxiter, o.history, o.gtol, o.ftol, o.fd_eps = int(maxiter), int(history), gtol, ftol, fd_eps
        n = x0.shape[0]
        x, f = np.empty((n, d)), np.empty(n)
        nit, nfev = ctypes.c_int(0), ctypes.c_longlong(0)
        fn = lib.carma_mle_batch_device if on_device else lib.carma_mle_batch
        check(fn(self.handle, kind, p, q, ctypes.byref(prior), flags, n, _ptr(x0), _ptr(lo), _ptr(hi),
                 ctypes.byref(o), _ptr(x), _ptr(f), ctypes.byref(nit), ctypes.byref(nfev), slot),
              "carma_mle_batch_device" if on_device else "carma_mle_batch")
        return x, f, nit.value, nfev.value

    def mle_grid(self, jobs, maxiter=1000, history=8, gtol=1e-5, ftol=2.2e-9, fd_eps=1e-8, slot=0):
        """Several models fitted in one launch (carma_mle_grid_device).  jobs: sequence of (kind, p, q, x0, lower, upper,
        prior, flags), the last six as CarmaModel.mle_starts returns them.  Returns a list of (x, f, nit, nfev), one per job."""
        nj = len(jobs)
        cj = (MLEJob * max(nj, 1))()
        lo = np.zeros((max(nj, 1), MAX_DIM))
        hi = np.zeros((max(nj, 1), MAX_DIM))
        xs, dims = [], []
        for j, (kind, p, q, x0, lower, upper, prior, flags) in enumerate(jobs):
            d = model_dim(kind, p, q)
            x0 = np.ascontiguousarray(x0, dtype=np.float64)
            if x0.ndim != 2 or x0.shape[1] != d:
                raise ValueError("x0 of job %d must be (nstart, %d)" % (j, d))
            lo[j, :d] = np.asarray(lower, dtype=np.float64).reshape(d)
            hi[j, :d] = np.asarray(upper, dtype=np.float64).reshape(d)
            cj[j].kind, cj[j].p, cj[j].q, cj[j].flags, cj[j].nstart = kind, p, q, flags, x0.shape[0]
            cj[j].prior = self.default_prior() if prior is None else prior
            xs.append(x0.reshape(-1))
            dims.append((x0.shape[0], d))
        xin = np.ascontiguousarray(np.concatenate(xs)) if xs else np.zeros(0)
        xout = np.empty_like(xin)
        fout = np.empty(sum(n for n, _ in dims))
        nit = (ctypes.c_int * max(nj, 1))()
        nfev = (ctypes.c_longlong * max(nj, 1))()
        o = MLEOpts()
        lib.carma_mle_default_opts(ctypes.byref(o))
        o.maxiter, o.history, o.gtol, o.ftol, o.fd_eps = maxiter, history, gtol, ftol, fd_eps
        check(lib.carma_mle_grid_device(self.handle, nj, cj, _ptr(xin), _ptr(lo), _ptr(hi), ctypes.byref(o), _ptr(xout),
                                        _ptr(fout), nit, nfev, slot), "carma_mle_grid_device")
        res, ox, of = [], 0, 0
        for j, (n, d) in enumerate(dims):
            res.append((xout[ox:ox + n * d].reshape(n, d).copy(), fout[of:of + n].copy(), int(nit[j]), int(nfev[j])))
            ox += n * d
            of += n
        return res

    def pt_run_dev(self, kind, p, q, opts, n_ensembles, d_samples, d_logposts, prior, d_init=None, d_accept=None,
                   d_exchange=None, stream=0):
        check(lib.carma_pt_run_dev(self.handle, kind, p, q, ctypes.byref(prior), ctypes.byref(opts), n_ensembles,
                                   d_init, d_samples, d_logposts, d_accept, d_exchange, stream), "carma_pt_run_dev")


class MultiSeries:
    """Ragged batch of light curves (CSR offsets) resident in HBM (carma_multi_series_t)."""

    def __init__(self, time, y, yerr, offsets, device=0):
        t, yy, ee = _c(time), _c(y), _c(yerr)
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        self.ncurves = off.size - 1
        self.handle = _vp()
        check(lib.carma_multi_series_create(_ptr(t), _ptr(yy), _ptr(ee),
                                            off.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), self.ncurves, device,
                                            ctypes.byref(self.handle)), "carma_multi_series_create")
        self.offsets = off
        self.device = device

    @classmethod
    def simulate(cls, ncurves, ny, kind, p, q, theta_true, yerr=0.05, dt_min=0.05, dt_max=50.0, seed=0,
                 curve_offset=0, prior=None, device=0):
        """ncurves synthetic light curves of ny points generated in HBM from the model theta_true
        (carma_multi_series_simulate); nothing crosses PCIe except the per-curve statistics."""
        th = _c(theta_true)
        if th.shape != (model_dim(kind, p, q),):
            raise ValueError("theta_true must have %d entries" % model_dim(kind, p, q))
        self = cls.__new__(cls)
        self.ncurves = int(ncurves)
        self.handle = _vp()
        pp = ctypes.byref(prior) if prior is not None else None
        check(lib.carma_multi_series_simulate(int(ncurves), int(ny), kind, p, q, _ptr(th), pp, float(yerr),
                                              float(dt_min), float(dt_max), int(seed), int(curve_offset), device,
                                              ctypes.byref(self.handle)), "carma_multi_series_simulate")
        self.offsets = np.arange(self.ncurves + 1, dtype=np.int64) * int(ny)
        self.device = device
        return self

    def curve(self, c):
        """(time, y, yerr) of curve c copied back to the host; time starts at 0."""
        n = int(self.offsets[c + 1] - self.offsets[c])
        t, y, e = np.empty(n), np.empty(n), np.empty(n)
        got = _sz(0)
        check(lib.carma_multi_series_get_curve(self.handle, int(c), _ptr(t), _ptr(y), _ptr(e), n, ctypes.byref(got)),
              "carma_multi_series_get_curve")
        return t, y, e

    def close(self):
        if getattr(self, "handle", None):
            lib.carma_multi_series_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def default_priors(self, population_var=True):
        out = np.empty(self.ncurves, dtype=PRIOR_DTYPE)
        check(lib.carma_multi_series_default_priors(self.handle, int(population_var), out.ctypes.data), "default_priors")
        return out

    def loglik(self, kind, p, q, theta, priors=None, flags=0):
        th = np.ascontiguousarray(theta, dtype=np.float64)
        d = model_dim(kind, p, q)
        if th.shape != (self.ncurves, d):
            raise ValueError("theta must be (%d, %d)" % (self.ncurves, d))
        out = np.empty(self.ncurves)
        pp = None
        if priors is not None:
            priors = np.ascontiguousarray(priors, dtype=PRIOR_DTYPE)
            pp = priors.ctypes.data
        check(lib.carma_multi_loglik(self.handle, kind, p, q, pp, th.ctypes.data, out.ctypes.data, flags),
              "carma_multi_loglik")
        return out

    def pt_run(self, kind, p, q, nsamples, burnin, thin=1, ntemps=10, n_ensembles=1, seed=1, ensemble_offset=0,
               priors=None, order_mode=0):
        """One PT-MCMC run per light curve (n_ensembles ensembles each), all curves in one launch."""
        d = model_dim(kind, p, q)
        o = PTOpts()
        lib.carma_pt_default_opts(ctypes.byref(o))
        o.nsamples, o.burnin, o.thin, o.ntemps = int(nsamples), int(burnin), int(thin), int(ntemps)
        o.seed, o.ensemble_offset, o.order_mode = seed, ensemble_offset, order_mode
        nc = self.ncurves
        samples = np.empty((nc, n_ensembles, nsamples, d))
        logposts = np.empty((nc, n_ensembles, nsamples))
        acc = np.empty((nc, n_ensembles, ntemps))
        xr = np.empty((nc, n_ensembles, ntemps))
        pp = None
        if priors is not None:
            priors = np.ascontiguousarray(priors, dtype=PRIOR_DTYPE)
            pp = priors.ctypes.data
        check(lib.carma_multi_pt_run(self.handle, kind, p, q, pp, ctypes.byref(o), n_ensembles, samples.ctypes.data,
                                     logposts.ctypes.data, acc.ctypes.data, xr.ctypes.data), "carma_multi_pt_run")
        return dict(samples=samples, logposts=logposts, accept_rates=acc, exchange_rates=xr)

    def loglik_dev(self, kind, p, q, d_priors_ptr, d_theta_ptr, d_out_ptr, flags=0, stream=0):
        check(lib.carma_multi_loglik_dev(self.handle, kind, p, q, d_priors_ptr, d_theta_ptr, d_out_ptr, flags, stream),
              "carma_multi_loglik_dev")


def lbfgs_batch(fun_batch, x0, lower, upper, maxiter=1000, history=8, gtol=1e-5, ftol=2.2e-9, fd_eps=1e-8):
    """The native optimiser core (carma_lbfgs_batch) on a Python objective: fun_batch maps an (n, d) array to n
    values.  Host code only; this is the loop Series.mle_batch runs with the GPU log-density as objective."""
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    n, d = x0.shape
    lo, hi = _c(np.broadcast_to(lower, (d,))), _c(np.broadcast_to(upper, (d,)))

    def trampoline(theta, npts, dim, fout, _user):
        try:
            th = np.ctypeslib.as_array(theta, shape=(npts, dim))
            np.ctypeslib.as_array(fout, shape=(npts,))[:] = np.asarray(fun_batch(th.copy()), dtype=np.float64)
            return 0
        except Exception:  # noqa: BLE001
            return 1

    cb = OBJECTIVE_FN(trampoline)
    o = MLEOpts()
    lib.carma_mle_default_opts(ctypes.byref(o))
    o.maxiter, o.history, o.gtol, o.ftol, o.fd_eps = int(maxiter), int(history), gtol, ftol, fd_eps
    x, f = np.empty((n, d)), np.empty(n)
    nit, nfev = ctypes.c_int(0), ctypes.c_longlong(0)
    check(lib.carma_lbfgs_batch(cb, None, d, n, _ptr(x0), _ptr(lo), _ptr(hi), ctypes.byref(o), _ptr(x), _ptr(f),
                                ctypes.byref(nit), ctypes.byref(nfev)), "carma_lbfgs_batch")
    return x, f, nit.value, nfev.value


def log_prior(kind, p, theta, prior):
    th = _c(theta)
    out = ctypes.c_double()
    check(lib.carma_log_prior(kind, p, _ptr(th), ctypes.byref(prior), ctypes.byref(out)), "carma_log_prior")
    return out.value


def fp64_peak_tflops(device=0):
    out = ctypes.c_double()
    check(lib.carma_fp64_peak_tflops(device, ctypes.byref(out)), "carma_fp64_peak_tflops")
    return out.value


def philox_dev(c0, c1, c2, c3, seed):
    out = (ctypes.c_uint32 * 4)()
    check(lib.carma_philox_dev(c0, c1, c2, c3, seed, out), "carma_philox_dev")
    return [int(x) for x in out]


def tdist_dev(seed, chain, it, j, dof=8):
    out = ctypes.c_double()
    check(lib.carma_tdist_dev(seed, chain, it, j, dof, ctypes.byref(out)), "carma_tdist_dev")
    return out.value


def fastmath_dev(rate, dt):
    """The time loop's transcendentals (csrc/fast_math.cuh) on arrays: rate in table steps per unit time.
    Returns exp(rate dt ln2/64), sin and cos(rate dt pi/64), (1-rho)/2 and (1+rho)/2 with rho = the exponential, 1/rate."""
    rate, dt = _c(rate), _c(dt)
    outs = [np.empty(rate.size) for _ in range(6)]
    check(lib.carma_fastmath_dev(_ptr(rate), _ptr(dt), rate.size, *[_ptr(o) for o in outs]), "carma_fastmath_dev")
    return tuple(outs)
