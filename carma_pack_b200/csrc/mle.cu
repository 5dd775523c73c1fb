// mle.cu -- maximum-likelihood fits from many random starts, all starts in lock-step (host code).
//
// Replaces the per-start scipy.optimize.minimize(method="L-BFGS-B") calls of the reference's
// _get_mle_single (src/carmcmc/carma_pack.py:195-252, objective _carma_loglik 255-260: one FFI crossing and one
// full filter run per function value, 2(d)+1 of them per finite-difference gradient).  Here every iteration of
// the optimiser is a handful of batched K1 launches over all starts: n*d perturbed points for the forward-
// difference gradients, n candidate points per backtracking step.  The host side of an iteration is O(n m d)
// flops in plain loops; it runs outside the Python interpreter, so fits of different (p,q) models driven from
// different host threads overlap on the GPU (each through its own series handle and stream slot).
//
// Algorithm: projected L-BFGS (two-loop recursion, history m) with forward-difference gradients (scipy's
// epsilon = 1e-8, stepping inward at an upper bound), Armijo backtracking along the projected path, stop on
// projected-gradient norm < gtol or relative decrease <= ftol (L-BFGS-B's factr*epsmch = 2.2e-9).  It is the
// C++ twin of carma_pack_b200.carma_pack.batched_lbfgs, which the tests keep as the cross-check.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "series.h"
#include "theta_transform.cuh"

using namespace carma;

extern "C" void carma_mle_default_opts(carma_mle_opts_t* o) {
    if (!o) return;
    o->maxiter = 1000;
    o->history = 8;
    o->max_backtrack = 25;
    o->reserved = 0;
    o->gtol = 1e-5;
    o->ftol = 2.2e-9;
    o->fd_eps = 1e-8;
}

namespace {

constexpr double BIG = 1e300;

struct PinnedBuf {
    double* p = nullptr;
    size_t cap = 0;
    bool reserve(size_t n) {
        if (n <= cap) return true;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        if (cudaHostAlloc((void**)&p, n * sizeof(double), cudaHostAllocDefault) != cudaSuccess) { p = nullptr; return false; }
        cap = n;
        return true;
    }
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
};

// what the optimiser minimises: a batch of points staged in `in` -> their function values
struct Objective {
    double* in = nullptr;   // staging buffer for the points of the next run(): capacity n (d + 4) d doubles
    long long nfev = 0;
    virtual int run(size_t n, double* f) = 0;  // f[k] for the first n staged points; non-finite -> BIG
    virtual ~Objective() {}
};

// -LogDensity(theta) through K1 (pinned buffers, the series' stream slot)
struct GpuObjective : Objective {
    carma_series_t s;
    int kind, p, q, slot;
    const carma_prior_t* prior;
    unsigned flags;
    PinnedBuf inbuf, out;
    int run(size_t n, double* f) override {
        if (n == 0) return CARMA_OK;
        int rc = carma_loglik_batch_async(s, kind, p, q, prior, n, inbuf.p, out.p, flags, slot);
        if (rc) return rc;
        rc = carma_loglik_batch_wait(s, slot);
        if (rc) return rc;
        for (size_t k = 0; k < n; k++) {
            double v = -out.p[k];
            f[k] = std::isfinite(v) ? v : BIG;
        }
        nfev += (long long)n;
        return CARMA_OK;
    }
};

// a caller-supplied objective (carma_lbfgs_batch): used by the CPU tests of the optimiser core
struct CallbackObjective : Objective {
    carma_objective_fn fn;
    void* user;
    size_t d;
    std::vector<double> buf;
    int run(size_t n, double* f) override {
        if (n == 0) return CARMA_OK;
        int rc = fn(in, n, d, f, user);
        if (rc) { set_error("carma_lbfgs_batch: the objective callback reported an error"); return CARMA_ERR_ARG; }
        for (size_t k = 0; k < n; k++)
            if (!std::isfinite(f[k])) f[k] = BIG;
        nfev += (long long)n;
        return CARMA_OK;
    }
};

int lbfgs_core(Objective& ev, size_t n, size_t d, const double* x0, const double* lower, const double* upper,
               const carma_mle_opts_t& o, double* x_out, double* f_out, int* nit_out, long long* nfev_out);

}  // namespace

static int check_opts(const carma_mle_opts_t* opts, carma_mle_opts_t& o, const char* who) {
    if (opts) o = *opts; else carma_mle_default_opts(&o);
    if (o.maxiter < 0 || o.history < 1 || o.history > 64 || o.max_backtrack < 1 || !(o.fd_eps > 0)) {
        set_error(std::string(who) + ": invalid options");
        return CARMA_ERR_ARG;
    }
    return CARMA_OK;
}

extern "C" int carma_mle_batch(carma_series_t s, int kind, int p, int q, const carma_prior_t* prior, unsigned flags,
                               size_t nstart, const double* x0, const double* lower, const double* upper,
                               const carma_mle_opts_t* opts, double* x_out, double* f_out, int* nit_out,
                               long long* nfev_out, int slot) {
    if (!s || !prior || !x0 || !lower || !upper || !x_out || !f_out || slot < 0 || slot >= CARMA_N_SLOTS) {
        set_error("carma_mle_batch: bad argument");
        return CARMA_ERR_ARG;
    }
    if (kind < CARMA_KIND_CAR1 || kind > CARMA_KIND_ZCARMA || p < 1 || p > MAX_P || (kind == CARMA_KIND_CAR1 && p != 1) ||
        (kind == CARMA_KIND_CARMA && !(q >= 0 && q < p))) {
        set_error("carma_mle_batch: invalid (kind,p,q)");
        return CARMA_ERR_ARG;
    }
    carma_mle_opts_t o;
    int rc = check_opts(opts, o, "carma_mle_batch");
    if (rc) return rc;
    if (nit_out) *nit_out = 0;
    if (nfev_out) *nfev_out = 0;
    if (nstart == 0) return CARMA_OK;
    if (!cuda_ok(cudaSetDevice(s->device), "cudaSetDevice")) return CARMA_ERR_CUDA;
    const size_t n = nstart, d = (size_t)model_dim(kind, p, q);
    GpuObjective ev;
    ev.s = s; ev.kind = kind; ev.p = p; ev.q = q; ev.slot = slot; ev.prior = prior; ev.flags = flags;
    if (!ev.inbuf.reserve(n * (d + 4) * d) || !ev.out.reserve(n * (d + 4))) { set_error("carma_mle_batch: pinned allocation failed"); return CARMA_ERR_ALLOC; }
    ev.in = ev.inbuf.p;
    return lbfgs_core(ev, n, d, x0, lower, upper, o, x_out, f_out, nit_out, nfev_out);
}

extern "C" int carma_lbfgs_batch(carma_objective_fn fn, void* user, size_t d, size_t nstart, const double* x0,
                                 const double* lower, const double* upper, const carma_mle_opts_t* opts, double* x_out,
                                 double* f_out, int* nit_out, long long* nfev_out) {
    if (!fn || !x0 || !lower || !upper || !x_out || !f_out || d == 0) { set_error("carma_lbfgs_batch: bad argument"); return CARMA_ERR_ARG; }
    carma_mle_opts_t o;
    int rc = check_opts(opts, o, "carma_lbfgs_batch");
    if (rc) return rc;
    if (nit_out) *nit_out = 0;
    if (nfev_out) *nfev_out = 0;
    if (nstart == 0) return CARMA_OK;
    CallbackObjective ev;
    ev.fn = fn; ev.user = user; ev.d = d;
    ev.buf.resize(nstart * (d + 4) * d);
    ev.in = ev.buf.data();
    return lbfgs_core(ev, nstart, d, x0, lower, upper, o, x_out, f_out, nit_out, nfev_out);
}

namespace {

int lbfgs_core(Objective& ev, size_t n, size_t d, const double* x0, const double* lower, const double* upper,
               const carma_mle_opts_t& o, double* x_out, double* f_out, int* nit_out, long long* nfev_out) {
    const int m = o.history;
    std::vector<double> x(n * d), f(n), g(n * d), xn(n * d), fn(n), gn(n * d), pg(n * d), qv(n * d), dir(n * d), slope(n), t(n);
    std::vector<double> S((size_t)m * n * d), Y((size_t)m * n * d), alpha((size_t)m * n), rho((size_t)m * n), ftmp(n * d);
    std::vector<char> active(n), todo(n), moved(n), blocked(n * d), grad_done(n);
    std::vector<double> fbig(n * (d + 4));
    std::vector<size_t> rows, retry;
    // History is kept PER ROW: row i has nh[i] pairs, pair h (oldest first) in ring slot (h0[i] + h) % m.  A row
    // whose newest (s, y) fails the curvature test simply keeps its previous pairs, so the result of a start never
    // depends on which other starts share the batch (and hence not on how the starts are sharded over GPUs).
    std::vector<int> nh(n, 0), h0(n, 0);
    // A row that stops on the small-decrease rule or on a failed line search first gets its history dropped and
    // continues from a scaled steepest-descent step (twice at most): with a poor quasi-Newton model one short Armijo
    // step can look like convergence far from a stationary point.
    std::vector<int> restarts_left(n, 2);

    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < d; j++) x[i * d + j] = std::min(std::max(x0[i * d + j], lower[j]), upper[j]);

    // forward-difference gradients of the listed rows at (z, fz) -> gout
    auto grad = [&](const std::vector<size_t>& rr, const std::vector<double>& z, const std::vector<double>& fz,
                    std::vector<double>& gout) -> int {
        size_t k = 0;
        for (size_t i : rr)
            for (size_t j = 0; j < d; j++, k++) {
                std::memcpy(ev.in + k * d, &z[i * d], d * sizeof(double));
                double h = (z[i * d + j] + o.fd_eps > upper[j]) ? -o.fd_eps : o.fd_eps;
                ev.in[k * d + j] += h;
            }
        int rc = ev.run(k, ftmp.data());
        if (rc) return rc;
        k = 0;
        retry.clear();
        for (size_t i : rr)
            for (size_t j = 0; j < d; j++, k++) {
                double h = (z[i * d + j] + o.fd_eps > upper[j]) ? -o.fd_eps : o.fd_eps;
                if (std::fabs(ftmp[k]) >= BIG) { gout[i * d + j] = 0.0; retry.push_back(i * d + j); }
                else gout[i * d + j] = (ftmp[k] - fz[i]) / h;
            }
        if (retry.empty()) return CARMA_OK;
        // the perturbed point has no finite value (it left the support of the density): difference the other way
        // before giving the component up, so that a start next to a prior bound is not mistaken for a stationary
        // point.  If the other side is infeasible too the component stays 0.
        k = 0;
        for (size_t ij : retry) {
            const size_t i = ij / d, j = ij % d;
            std::memcpy(ev.in + k * d, &z[i * d], d * sizeof(double));
            const double h = (z[i * d + j] + o.fd_eps > upper[j]) ? -o.fd_eps : o.fd_eps;
            const double hb = -h;
            if (z[i * d + j] + hb >= lower[j] && z[i * d + j] + hb <= upper[j]) ev.in[k * d + j] += hb;
            k++;
        }
        rc = ev.run(k, ftmp.data());
        if (rc) return rc;
        k = 0;
        for (size_t ij : retry) {
            const size_t i = ij / d, j = ij % d;
            const double h = (z[i * d + j] + o.fd_eps > upper[j]) ? -o.fd_eps : o.fd_eps;
            const double hb = -h;
            const bool stepped = z[i * d + j] + hb >= lower[j] && z[i * d + j] + hb <= upper[j];
            if (stepped && std::fabs(ftmp[k]) < BIG) gout[ij] = (ftmp[k] - fz[i]) / hb;
            k++;
        }
        return CARMA_OK;
    };

    std::memcpy(ev.in, x.data(), n * d * sizeof(double));
    int rc = ev.run(n, f.data());
    if (rc) return rc;
    rows.resize(n);
    for (size_t i = 0; i < n; i++) rows[i] = i;
    rc = grad(rows, x, f, g);
    if (rc) return rc;
    for (size_t i = 0; i < n; i++) active[i] = f[i] < BIG;

    int nit = 0;
    for (nit = 1; nit <= o.maxiter; nit++) {
        // projected gradient: zero the components pushing against an active bound
        bool any_active = false;
        for (size_t i = 0; i < n; i++) {
            double gmax = 0.0;
            for (size_t j = 0; j < d; j++) {
                const double xi = x[i * d + j], gi = g[i * d + j];
                const bool blk = (xi <= lower[j] && gi > 0) || (xi >= upper[j] && gi < 0);
                blocked[i * d + j] = blk;
                pg[i * d + j] = blk ? 0.0 : gi;
                gmax = std::max(gmax, std::fabs(pg[i * d + j]));
            }
            if (gmax < o.gtol) active[i] = 0;
            any_active = any_active || active[i];
        }
        if (!any_active) break;
        // two-loop recursion, row by row
        for (size_t i = 0; i < n; i++) {
            double* qi = &qv[i * d];
            const double* pgi = &pg[i * d];
            for (size_t j = 0; j < d; j++) qi[j] = pgi[j];
            const int nhist = nh[i], hist0 = h0[i];
            for (int h = nhist - 1; h >= 0; h--) {
                const size_t sl = (size_t)((hist0 + h) % m);
                const double *sv = &S[(sl * n + i) * d], *yv = &Y[(sl * n + i) * d];
                double sy = 0.0, sq = 0.0;
                for (size_t j = 0; j < d; j++) { sy += sv[j] * yv[j]; sq += sv[j] * qi[j]; }
                const double r = 1.0 / std::max(sy, 1e-300), a = r * sq;
                rho[sl * n + i] = r;
                alpha[sl * n + i] = a;
                for (size_t j = 0; j < d; j++) qi[j] -= a * yv[j];
            }
            if (nhist > 0) {
                const size_t sl = (size_t)((hist0 + nhist - 1) % m);
                const double *sv = &S[(sl * n + i) * d], *yv = &Y[(sl * n + i) * d];
                double sy = 0.0, yy = 0.0;
                for (size_t j = 0; j < d; j++) { sy += sv[j] * yv[j]; yy += yv[j] * yv[j]; }
                const double gam = std::min(std::max(sy / std::max(yy, 1e-300), 1e-8), 1e8);
                for (size_t j = 0; j < d; j++) qi[j] *= gam;
            } else {
                double nrm = 0.0;
                for (size_t j = 0; j < d; j++) nrm += pgi[j] * pgi[j];
                const double sc = 1.0 / std::max(std::sqrt(nrm), 1.0);
                for (size_t j = 0; j < d; j++) qi[j] *= sc;
            }
            for (int h = 0; h < nhist; h++) {
                const size_t sl = (size_t)((hist0 + h) % m);
                const double *sv = &S[(sl * n + i) * d], *yv = &Y[(sl * n + i) * d];
                double yq = 0.0;
                for (size_t j = 0; j < d; j++) yq += yv[j] * qi[j];
                const double b = rho[sl * n + i] * yq, a = alpha[sl * n + i];
                for (size_t j = 0; j < d; j++) qi[j] += (a - b) * sv[j];
            }
            double sl_ = 0.0, pg2 = 0.0;
            for (size_t j = 0; j < d; j++) {
                dir[i * d + j] = blocked[i * d + j] ? 0.0 : -qi[j];
                sl_ += dir[i * d + j] * pgi[j];
                pg2 += pgi[j] * pgi[j];
            }
            if (!(sl_ < 0)) {
                for (size_t j = 0; j < d; j++) dir[i * d + j] = -pgi[j];
                sl_ = -pg2;
            }
            slope[i] = sl_;
        }
        // batched Armijo backtracking on the projected path.  The step sizes are tried in the usual order
        // 1, 1/2, 1/4, ... and the first one that satisfies the condition is taken, but FOUR consecutive sizes of
        // every row still searching are evaluated per launch, and the first launch also carries the d finite-
        // difference points around the full-step candidate: when the full step is accepted (the common case) the
        // next gradient is already there.  Same iterates as one-size-at-a-time backtracking, ~3x fewer launches.
        for (size_t i = 0; i < n; i++) { t[i] = 1.0; todo[i] = active[i]; fn[i] = f[i]; grad_done[i] = 0; }
        xn = x;
        gn = g;
        int tried = 0;
        for (int round = 0; tried < o.max_backtrack; round++) {
            rows.clear();
            for (size_t i = 0; i < n; i++) if (todo[i]) rows.push_back(i);
            if (rows.empty()) break;
            const int nt = std::min(4, o.max_backtrack - tried);
            const bool spec = (round == 0);
            const size_t per_row = (size_t)nt + (spec ? d : 0);
            size_t k = 0;
            for (size_t i : rows) {
                double* base = ev.in + k * per_row * d;
                double tk = t[i];
                for (int c = 0; c < nt; c++, tk *= 0.5)
                    for (size_t j = 0; j < d; j++)
                        base[(size_t)c * d + j] = std::min(std::max(x[i * d + j] + tk * dir[i * d + j], lower[j]), upper[j]);
                if (spec)
                    for (size_t j = 0; j < d; j++) {
                        double* pt = base + ((size_t)nt + j) * d;
                        std::memcpy(pt, base, d * sizeof(double));
                        pt[j] += (base[j] + o.fd_eps > upper[j]) ? -o.fd_eps : o.fd_eps;
                    }
                k++;
            }
            rc = ev.run(k * per_row, fbig.data());
            if (rc) return rc;
            k = 0;
            for (size_t i : rows) {
                const double* fr = &fbig[k * per_row];
                const double* base = ev.in + k * per_row * d;
                double tk = t[i];
                int hit = -1;
                for (int c = 0; c < nt; c++, tk *= 0.5)
                    if (fr[c] <= f[i] + 1e-4 * tk * slope[i]) { hit = c; break; }
                if (hit >= 0) {
                    std::memcpy(&xn[i * d], base + (size_t)hit * d, d * sizeof(double));
                    fn[i] = fr[hit];
                    todo[i] = 0;
                    if (spec && hit == 0) {
                        bool all_finite = true;
                        for (size_t j = 0; j < d; j++) {
                            const double h = (base[j] + o.fd_eps > upper[j]) ? -o.fd_eps : o.fd_eps;
                            const double fv = fr[(size_t)nt + j];
                            if (std::fabs(fv) >= BIG) all_finite = false;
                            gn[i * d + j] = (fv - fn[i]) / h;
                        }
                        grad_done[i] = all_finite;  // otherwise grad() below redoes the row with its backward retry
                    }
                } else {
                    for (int c = 0; c < nt; c++) t[i] *= 0.5;
                }
                k++;
            }
            tried += nt;
        }
        rows.clear();
        for (size_t i = 0; i < n; i++) {
            moved[i] = active[i] && !todo[i];
            if (moved[i] && !grad_done[i]) rows.push_back(i);
        }
        if (!rows.empty()) {
            rc = grad(rows, xn, fn, gn);
            if (rc) return rc;
        }
        // history update, row by row: a pair enters a row's ring only if that row moved and s.y > 0
        for (size_t i = 0; i < n; i++) {
            if (!moved[i]) continue;
            double sy = 0.0;
            for (size_t j = 0; j < d; j++) sy += (xn[i * d + j] - x[i * d + j]) * (gn[i * d + j] - g[i * d + j]);
            if (!(sy > 1e-12)) continue;
            size_t dst;
            if (nh[i] == m) {  // drop the oldest
                dst = (size_t)h0[i];
                h0[i] = (h0[i] + 1) % m;
            } else {
                dst = (size_t)((h0[i] + nh[i]) % m);
                nh[i]++;
            }
            double *sv = &S[(dst * n + i) * d], *yv = &Y[(dst * n + i) * d];
            for (size_t j = 0; j < d; j++) { sv[j] = xn[i * d + j] - x[i * d + j]; yv[j] = gn[i * d + j] - g[i * d + j]; }
        }
        for (size_t i = 0; i < n; i++) {
            const bool small = moved[i] && ((f[i] - fn[i]) <= o.ftol * std::max(std::max(std::fabs(f[i]), std::fabs(fn[i])), 1.0));
            if (todo[i] || small) {  // line search failed / negligible decrease
                if (nh[i] > 0 && restarts_left[i] > 0) {
                    nh[i] = 0;
                    h0[i] = 0;
                    restarts_left[i]--;
                } else {
                    active[i] = 0;
                }
            }
        }
        x = xn;
        f = fn;
        g = gn;
    }
    std::memcpy(x_out, x.data(), n * d * sizeof(double));
    std::memcpy(f_out, f.data(), n * sizeof(double));
    if (nit_out) *nit_out = std::min(nit, o.maxiter);
    if (nfev_out) *nfev_out = ev.nfev;
    return CARMA_OK;
}

}  // namespace
