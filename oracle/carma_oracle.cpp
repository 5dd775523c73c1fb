// carma_oracle.cpp -- CPU ORACLE for the carma_pack hot path.  TEST INFRASTRUCTURE ONLY.
//
// This file is a dependency-free C++17 restatement of the reference algorithm
// (brandonckelly/carma_pack).  It is NOT part of the product: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load it.  The product path (carma_pack_b200/csrc) never includes or links it.
//
// Parity pin: validated against the reference's own numpy restatement
// (src/carmcmc/carma_pack.py:1264-1488 KalmanFilterDeprecated, 1084-1123
// carma_variance) run in the build container, and against the reference test-suite's
// known answers (cpp_tests/carma_unit_tests.cpp:441-444, 1313-1316).  See
// tests/golden/make_golden.py and tests/test_oracle_golden.py.
//
// The reference itself cannot be compiled here (needs Armadillo + Boost, absent,
// no network), so the dense complex algebra that Armadillo/LAPACK would do is
// restated explicitly: arma::solve -> LU with partial pivoting (zgesv semantics),
// symmatu -> Hermitian mirror, arma::exp/pow -> std::exp/std::pow.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/src).
//
// The arithmetic is templated on the real type so that the same restatement can
// be evaluated in long double: |double - long double| is the intrinsic rounding
// noise of the reference algorithm at a given theta, which the parity tests use
// to recognise ill-conditioned parameter vectors.

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace {

template <class R> using cx = std::complex<R>;

enum ModelKind { KIND_CAR1 = 0, KIND_CARP = 1, KIND_CARMA = 2, KIND_ZCAR = 3, KIND_ZCARMA = 4 };

struct Prior {
    double max_stdev;   // carpack.hpp:201-207 SetPrior
    double max_freq;    // 1 / min(dt)
    double min_freq;    // 1 / (tmax - tmin)
    double kappa_low;   // carpack.hpp:413-419 (ZCARMA only)
    double kappa_high;
    double measerr_dof; // carpack.hpp:63 (=50)
};

template <class R> R pi_v() { return std::acos(R(-1)); }

// ---------------------------------------------------------------------------------
// carpack.cpp:137-172  CARp::ARRoots  (also the MA-root recipe of carpack.cpp:526-555)
// ---------------------------------------------------------------------------------
template <class R>
void quad_roots(const R* logq, int n, std::vector<cx<R>>& roots) {
    roots.assign(n, cx<R>(0, 0));
    for (int i = 0; i < n / 2; i++) {
        R quad_term1 = std::exp(logq[2 * i]);
        R quad_term2 = std::exp(logq[2 * i + 1]);
        R discriminant = quad_term2 * quad_term2 - R(4.0) * quad_term1;
        if (discriminant > 0) {
            R root1 = R(-0.5) * (quad_term2 + std::sqrt(discriminant));
            R root2 = R(-0.5) * (quad_term2 - std::sqrt(discriminant));
            roots[2 * i] = cx<R>(root1, 0);
            roots[2 * i + 1] = cx<R>(root2, 0);
        } else {
            R real_part = R(-0.5) * quad_term2;
            R imag_part = R(-0.5) * std::sqrt(-discriminant);
            roots[2 * i] = cx<R>(real_part, imag_part);
            roots[2 * i + 1] = cx<R>(real_part, -imag_part);
        }
    }
    if (n % 2 == 1) {
        R real_root = -std::exp(logq[n - 1]);
        roots[n - 1] = cx<R>(real_root, 0);
    }
}

// carpack.cpp:742-756 polycoefs
template <class R>
std::vector<R> polycoefs(const std::vector<cx<R>>& roots) {
    size_t n = roots.size();
    std::vector<cx<R>> coefs(n + 1, cx<R>(0, 0));
    coefs[0] = cx<R>(1, 0);
    for (size_t i = 0; i < n; i++) {
        // coefs(1..i+1) = coefs(1..i+1) - roots(i) * coefs(0..i), evaluated on the OLD values
        std::vector<cx<R>> old(coefs);
        for (size_t j = 1; j <= i + 1; j++) coefs[j] = old[j] - roots[i] * old[j - 1];
    }
    std::vector<R> out(n + 1);
    for (size_t j = 0; j <= n; j++) out[j] = coefs[j].real();
    return out;
}

// boost::math::binomial_coefficient<double>(n,k) (carpack.hpp:356, carpack.cpp:695)
template <class R> R binom(int n, int k) {
    R r = 1;
    for (int i = 1; i <= k; i++) r = r * R(n - k + i) / R(i);
    return std::round(r);
}

// carpack.cpp:704-705
template <class R> R logit(R x) { return std::log(x / (R(1) - x)); }
template <class R> R inv_logit(R x) { return std::exp(x) / (R(1) + std::exp(x)); }

// carpack.cpp:522-580 CARMA::ExtractMA ; carpack.cpp:687-698 ZCARMA::ExtractMA ;
// carpack.hpp:292-295 CARp ma_coefs_ = [1,0,...]; ZCAR shadowing quirk.
template <class R>
std::vector<R> extract_ma(int kind, const R* theta, int p, int q, const Prior& pr) {
    std::vector<R> ma(p, R(0));
    if (kind == KIND_CARMA && q > 0) {
        std::vector<cx<R>> ma_roots;
        quad_roots<R>(theta + 3 + p, q, ma_roots);
        std::vector<R> poly = polycoefs<R>(ma_roots);
        R norm = poly[q];
        for (int i = 0; i <= q; i++) poly[i] = poly[i] / norm;
        for (int i = 0; i < q + 1; i++) ma[i] = poly[q - i];
    } else if (kind == KIND_ZCARMA) {
        R kappa_normed = inv_logit<R>(theta[3 + p]);
        R kappa = (R(pr.kappa_high) - R(pr.kappa_low)) * kappa_normed + R(pr.kappa_low);
        ma[0] = 1;
        for (int i = 1; i < p; i++) ma[i] = binom<R>(p - 1, i) / std::pow(kappa, R(i));
    } else {
        ma[0] = 1;  // CAR1 / CARp / ZCAR(as evaluated by the reference)
    }
    return ma;
}

// carpack.cpp:377-409 CARp::Variance
template <class R>
R carma_variance(const std::vector<cx<R>>& roots, const std::vector<R>& ma, R sigma, R dt) {
    cx<R> car_var(0, 0);
    size_t p = roots.size();
    for (size_t k = 0; k < p; k++) {
        cx<R> denom_product(1, 0);
        for (size_t l = 0; l < p; l++) {
            if (l != k) denom_product *= (roots[l] - roots[k]) * (std::conj(roots[l]) + roots[k]);
        }
        cx<R> denom = R(-2.0) * roots[k].real() * denom_product;
        cx<R> ma_sum1(0, 0), ma_sum2(0, 0);
        for (size_t l = 0; l < ma.size(); l++) {
            ma_sum1 += ma[l] * std::pow(roots[k], R(l));
            ma_sum2 += ma[l] * std::pow(-roots[k], R(l));
        }
        cx<R> numer = ma_sum1 * ma_sum2 * std::exp(roots[k] * dt);
        car_var += numer / denom;
    }
    return sigma * sigma * car_var.real();
}

// carpack.cpp:709-732 unique_roots
template <class R>
bool unique_roots(const std::vector<cx<R>>& roots, R tolerance) {
    R min_frac_diff = R(100.0) * tolerance;
    int p = (int)roots.size();
    for (int i = 0; i < p - 1; i++)
        for (int j = i + 1; j < p; j++) {
            R frac_diff = std::abs((roots[i] - roots[j]) / (roots[i] + roots[j]));
            if (frac_diff < min_frac_diff) min_frac_diff = frac_diff;
        }
    return min_frac_diff > tolerance;
}

// carpack.hpp:178-191 (base), carpack.cpp:116-130 (CAR1), carpack.cpp:314-374 (CARp)
template <class R>
bool check_prior_bounds(int kind, const R* theta, int p, const Prior& pr, bool ignore_prior) {
    R ysigma = theta[0], measerr_scale = theta[1];
    if (kind == KIND_CAR1) {
        // CAR1::CheckPriorBounds does not look at ignore_prior_ (carpack.cpp:116)
        R omega = std::exp(theta[3]);
        if ((omega > R(pr.max_freq)) || (omega < R(pr.min_freq)) || (ysigma > R(pr.max_stdev)) ||
            (ysigma < 0) || (measerr_scale < R(0.5)) || (measerr_scale > R(2.0)))
            return false;
        return true;
    }
    if (ignore_prior) return true;
    std::vector<cx<R>> ar_roots;
    quad_roots<R>(theta + 3, p, ar_roots);
    R two_pi = R(2.0) * pi_v<R>();
    int ok1 = 0, ok2 = 0, ok3 = 0;
    std::vector<R> cent(p), width(p);
    for (int i = 0; i < p; i++) {
        cent[i] = std::abs(ar_roots[i].imag()) / R(2.0) / pi_v<R>();
        width[i] = -ar_roots[i].real() / R(2.0) / pi_v<R>();
        if (cent[i] < R(pr.max_freq)) ok1++;
        if (width[i] < R(pr.max_freq)) ok2++;
        if (width[i] > R(pr.min_freq)) ok3++;
    }
    (void)two_pi;
    bool prior_satisfied = unique_roots<R>(ar_roots, R(1e-4));
    if ((ok1 != p) || (ok2 != p) || (ok3 != p) || (ysigma > R(pr.max_stdev)) || (ysigma < 0) ||
        (measerr_scale < R(0.5)) || (measerr_scale > R(2.0)))
        prior_satisfied = false;
    for (int i = 1; i < p; i++) {  // order_lorentzians_ == true (carpack.hpp:294)
        R d = cent[i] - cent[i - 1];
        if (d > R(1e-8)) prior_satisfied = false;
    }
    return prior_satisfied;
}

// carpack.hpp:118-126 (base) and carpack.hpp:444-456 (ZCARMA)
template <class R>
R log_prior(int kind, const R* theta, int p, const Prior& pr) {
    R measerr_scale = theta[1];
    R dof = R(pr.measerr_dof);
    R logprior = R(-0.5) * dof / measerr_scale - (R(1.0) + dof / R(2.0)) * std::log(measerr_scale);
    if (kind == KIND_ZCARMA) {
        R logit_kappa = theta[p + 3];
        logprior += -logit_kappa - R(2.0) * std::log(R(1.0) + std::exp(-logit_kappa));
    }
    return logprior;
}

// ---------------------------------------------------------------------------------
// Dense complex helpers standing in for Armadillo
// ---------------------------------------------------------------------------------
// arma::solve(A, b) for a square system -> LAPACK zgesv: LU with partial pivoting,
// pivot chosen by |re|+|im| (izamax).  Returns false when a pivot is exactly zero
// (zgesv info>0 -> arma::solve throws std::runtime_error, carpack.hpp:154-164).
template <class R>
bool lu_solve(std::vector<cx<R>> A, int n, std::vector<cx<R>>& b) {
    auto at = [&](int i, int j) -> cx<R>& { return A[(size_t)i * n + j]; };
    for (int k = 0; k < n; k++) {
        int piv = k;
        R best = std::abs(at(k, k).real()) + std::abs(at(k, k).imag());
        for (int i = k + 1; i < n; i++) {
            R v = std::abs(at(i, k).real()) + std::abs(at(i, k).imag());
            if (v > best) { best = v; piv = i; }
        }
        if (!(best > 0) || !std::isfinite((double)best)) return false;
        if (piv != k) {
            for (int j = 0; j < n; j++) std::swap(at(k, j), at(piv, j));
            std::swap(b[k], b[piv]);
        }
        for (int i = k + 1; i < n; i++) {
            cx<R> l = at(i, k) / at(k, k);
            at(i, k) = l;
            for (int j = k + 1; j < n; j++) at(i, j) -= l * at(k, j);
            b[i] -= l * b[k];
        }
    }
    for (int i = n - 1; i >= 0; i--) {
        cx<R> s = b[i];
        for (int j = i + 1; j < n; j++) s -= at(i, j) * b[j];
        b[i] = s / at(i, i);
    }
    return true;
}

// ---------------------------------------------------------------------------------
// KalmanFilterp  (kfilter.hpp:266-389, kfilter.cpp:138-337) -- dense, as written
// ---------------------------------------------------------------------------------
template <class R>
struct KFp {
    int p = 0;
    size_t ny = 0;
    std::vector<R> time, y, yerr, dt;  // y is already centred; yerr already scaled
    R sigsqr = 0;
    std::vector<cx<R>> omega;
    std::vector<R> ma;
    std::vector<R> mean, var;
    // state
    std::vector<cx<R>> x, b, K, rho, state_const, state_slope;
    std::vector<cx<R>> V, P;  // p x p, row-major
    R innovation = 0;
    size_t cur = 0;
    R yconst = 0, yslope = 0;

    cx<R>& Vij(int i, int j) { return V[(size_t)i * p + j]; }
    cx<R>& Pij(int i, int j) { return P[(size_t)i * p + j]; }

    void set_series(const R* t, const R* yy, const R* ye, size_t n) {
        ny = n;
        time.assign(t, t + n); y.assign(yy, yy + n); yerr.assign(ye, ye + n);
        dt.resize(n > 0 ? n - 1 : 0);
        for (size_t i = 0; i + 1 < n; i++) dt[i] = time[i + 1] - time[i];  // kfilter.hpp:45
        mean.assign(n, 0); var.assign(n, 0);
    }
    void set_params(R s2, const std::vector<cx<R>>& om, const std::vector<R>& m) {
        sigsqr = s2; omega = om; p = (int)om.size();
        ma = m; ma.resize(p, R(0));  // kfilter.hpp:310-312
        x.assign(p, 0); b.assign(p, 0); K.assign(p, 0); rho.assign(p, 0);
        state_const.assign(p, 0); state_slope.assign(p, 0);
        V.assign((size_t)p * p, 0); P.assign((size_t)p * p, 0);
    }

    // Re( b * M * b^H )  (kfilter.cpp:181, 209)
    R quad_form(const std::vector<cx<R>>& M) {
        cx<R> tot(0, 0);
        for (int j = 0; j < p; j++) {
            cx<R> bM(0, 0);
            for (int i = 0; i < p; i++) bM += b[i] * M[(size_t)i * p + j];
            tot += bM * std::conj(b[j]);
        }
        return tot.real();
    }

    // kfilter.cpp:138-186
    bool Reset() {
        std::vector<cx<R>> E((size_t)p * p);
        for (int k = 0; k < p; k++) E[k] = cx<R>(1, 0);
        if (p > 1) for (int k = 0; k < p; k++) E[(size_t)p + k] = omega[k];
        for (int i = 2; i < p; i++)
            for (int k = 0; k < p; k++) E[(size_t)i * p + k] = std::pow(omega[k], R(i));
        std::vector<cx<R>> J(p, cx<R>(0, 0));
        J[p - 1] = cx<R>(1, 0);
        if (!lu_solve<R>(E, p, J)) return false;
        for (int k = 0; k < p; k++) {
            cx<R> s(0, 0);
            for (int i = 0; i < p; i++) s += ma[i] * E[(size_t)i * p + k];
            b[k] = s;
        }
        for (int i = 0; i < p; i++)
            for (int j = i; j < p; j++) {
                Vij(i, j) = -sigsqr * J[i] * std::conj(J[j]) / (omega[i] + std::conj(omega[j]));
            }
        for (int i = 0; i < p; i++)
            for (int j = 0; j < i; j++) Vij(i, j) = std::conj(Vij(j, i));  // symmatu (Hermitian, SURVEY 8a note)
        P = V;
        std::fill(x.begin(), x.end(), cx<R>(0, 0));
        mean[0] = 0;
        var[0] = quad_form(V) + yerr[0] * yerr[0];
        innovation = y[0];
        cur = 1;
        return true;
    }

    // shared by Update (kfilter.cpp:191-204) and Predict (kfilter.cpp:243-249)
    void gain_and_advance(R var_prev, R dtt) {
        for (int i = 0; i < p; i++) {
            cx<R> s(0, 0);
            for (int j = 0; j < p; j++) s += Pij(i, j) * std::conj(b[j]);
            K[i] = s / var_prev;
        }
        for (int i = 0; i < p; i++) x[i] += K[i] * innovation;
        for (int i = 0; i < p; i++)
            for (int j = 0; j < p; j++) Pij(i, j) -= var_prev * (K[i] * std::conj(K[j]));
        for (int i = 0; i < p; i++) rho[i] = std::exp(omega[i] * dtt);
        for (int i = 0; i < p; i++) x[i] = rho[i] * x[i];
        for (int i = 0; i < p; i++)
            for (int j = 0; j < p; j++)
                Pij(i, j) = (rho[i] * std::conj(rho[j])) * (Pij(i, j) - Vij(i, j)) + Vij(i, j);
    }

    // kfilter.cpp:189-215
    void Update() {
        gain_and_advance(var[cur - 1], dt[cur - 1]);
        cx<R> m(0, 0);
        for (int i = 0; i < p; i++) m += b[i] * x[i];
        mean[cur] = m.real();
        var[cur] = quad_form(P);
        var[cur] += yerr[cur] * yerr[cur];
        innovation = y[cur] - mean[cur];
        cur++;
    }

    // kfilter.hpp:126-132
    bool Filter() {
        if (!Reset()) return false;
        for (size_t i = 1; i < ny; i++) Update();
        return true;
    }

    // kfilter.cpp:290-312
    void InitializeCoefs(R tq, size_t itime, R ymean, R yvar) {
        for (int i = 0; i < p; i++) {
            cx<R> s(0, 0);
            for (int j = 0; j < p; j++) s += Pij(i, j) * std::conj(b[j]);
            K[i] = s / yvar;
        }
        for (int i = 0; i < p; i++) { state_const[i] = x[i] - K[i] * ymean; state_slope[i] = K[i]; }
        for (int i = 0; i < p; i++)
            for (int j = 0; j < p; j++) Pij(i, j) -= yvar * (K[i] * std::conj(K[j]));
        R dtt = std::abs(time[itime] - tq);
        for (int i = 0; i < p; i++) rho[i] = std::exp(omega[i] * dtt);
        for (int i = 0; i < p; i++) { state_const[i] = rho[i] * state_const[i]; state_slope[i] = rho[i] * state_slope[i]; }
        for (int i = 0; i < p; i++)
            for (int j = 0; j < p; j++)
                Pij(i, j) = (rho[i] * std::conj(rho[j])) * (Pij(i, j) - Vij(i, j)) + Vij(i, j);
        cx<R> c(0, 0), s(0, 0);
        for (int i = 0; i < p; i++) { c += b[i] * state_const[i]; s += b[i] * state_slope[i]; }
        yconst = c.real(); yslope = s.real();
        var[itime] = quad_form(P) + yerr[itime] * yerr[itime];
        cur = itime + 1;
    }

    // kfilter.cpp:316-337
    void UpdateCoefs() {
        R vprev = var[cur - 1];
        for (int i = 0; i < p; i++) {
            cx<R> s(0, 0);
            for (int j = 0; j < p; j++) s += Pij(i, j) * std::conj(b[j]);
            K[i] = s / vprev;
        }
        for (int i = 0; i < p; i++) {
            state_const[i] += K[i] * (y[cur - 1] - yconst);
            state_slope[i] -= K[i] * yslope;
        }
        for (int i = 0; i < p; i++)
            for (int j = 0; j < p; j++) Pij(i, j) -= vprev * (K[i] * std::conj(K[j]));
        for (int i = 0; i < p; i++) rho[i] = std::exp(omega[i] * dt[cur - 1]);
        for (int i = 0; i < p; i++) { state_const[i] = rho[i] * state_const[i]; state_slope[i] = rho[i] * state_slope[i]; }
        for (int i = 0; i < p; i++)
            for (int j = 0; j < p; j++)
                Pij(i, j) = (rho[i] * std::conj(rho[j])) * (Pij(i, j) - Vij(i, j)) + Vij(i, j);
        cx<R> c(0, 0), s(0, 0);
        for (int i = 0; i < p; i++) { c += b[i] * state_const[i]; s += b[i] * state_slope[i]; }
        yconst = c.real(); yslope = s.real();
        var[cur] = quad_form(P) + yerr[cur] * yerr[cur];
        cur++;
    }

    // kfilter.cpp:218-286
    bool Predict(R tq, R* out_mean, R* out_var) {
        size_t ipredict = 0;
        while (tq > time[ipredict]) {
            ipredict++;
            if (ipredict == ny) break;
        }
        if (!Reset()) return false;
        for (size_t i = 1; i < ipredict; i++) Update();
        R ypredict_mean, ypredict_var, yprecision;
        if (ipredict == 0) {
            ypredict_mean = 0;
            ypredict_var = quad_form(V);
        } else {
            gain_and_advance(var[ipredict - 1], std::abs(tq - time[ipredict - 1]));
            cx<R> m(0, 0);
            for (int i = 0; i < p; i++) m += b[i] * x[i];
            ypredict_mean = m.real();
            ypredict_var = quad_form(P);
        }
        if (ipredict == ny) { *out_mean = ypredict_mean; *out_var = ypredict_var; return true; }
        yprecision = R(1.0) / ypredict_var;
        ypredict_mean *= yprecision;
        InitializeCoefs(tq, ipredict, ypredict_mean / yprecision, ypredict_var);
        yprecision += yslope * yslope / var[ipredict];
        ypredict_mean += yslope * (y[ipredict] - yconst) / var[ipredict];
        for (size_t i = ipredict + 1; i < ny; i++) {
            UpdateCoefs();
            yprecision += yslope * yslope / var[i];
            ypredict_mean += yslope * (y[i] - yconst) / var[i];
        }
        ypredict_var = R(1.0) / yprecision;
        ypredict_mean *= ypredict_var;
        *out_mean = ypredict_mean; *out_var = ypredict_var;
        return true;
    }
};

// ---------------------------------------------------------------------------------
// KalmanFilter1 (kfilter.cpp:19-135)
// ---------------------------------------------------------------------------------
template <class R>
struct KF1 {
    size_t ny = 0;
    std::vector<R> time, y, yerr, dt, mean, var;
    R sigsqr = 0, omega = 0, yconst = 0, yslope = 0;
    size_t cur = 0;
    void set_series(const R* t, const R* yy, const R* ye, size_t n) {
        ny = n; time.assign(t, t + n); y.assign(yy, yy + n); yerr.assign(ye, ye + n);
        dt.resize(n > 0 ? n - 1 : 0);
        for (size_t i = 0; i + 1 < n; i++) dt[i] = time[i + 1] - time[i];
        mean.assign(n, 0); var.assign(n, 0);
    }
    void Reset() {  // kfilter.cpp:19-26
        mean[0] = 0; var[0] = sigsqr / (R(2.0) * omega) + yerr[0] * yerr[0];
        yconst = 0; yslope = 0; cur = 1;
    }
    void Update() {  // kfilter.cpp:29-48
        R rho = std::exp(R(-1.0) * omega * dt[cur - 1]);
        R previous_var = var[cur - 1] - yerr[cur - 1] * yerr[cur - 1];
        R var_ratio = previous_var / var[cur - 1];
        mean[cur] = rho * mean[cur - 1] + rho * var_ratio * (y[cur - 1] - mean[cur - 1]);
        var[cur] = sigsqr / (R(2.0) * omega) * (R(1.0) - rho * rho) + rho * rho * previous_var * (R(1.0) - var_ratio);
        var[cur] += yerr[cur] * yerr[cur];
        cur++;
    }
    void Filter() { Reset(); for (size_t i = 1; i < ny; i++) Update(); }
    void InitializeCoefs(R tq, size_t itime) {  // kfilter.cpp:51-56
        yconst = 0;
        yslope = std::exp(-std::abs(time[itime] - tq) * omega);
        var[itime] = sigsqr / (R(2.0) * omega) * (R(1.0) - yslope * yslope) + yerr[itime] * yerr[itime];
        cur = itime + 1;
    }
    void UpdateCoefs() {  // kfilter.cpp:59-69
        R rho = std::exp(R(-1.0) * dt[cur - 1] * omega);
        R previous_var = var[cur - 1] - yerr[cur - 1] * yerr[cur - 1];
        R var_ratio = previous_var / var[cur - 1];
        yslope *= rho * (R(1.0) - var_ratio);
        yconst = yconst * rho * (R(1.0) - var_ratio) + rho * var_ratio * y[cur - 1];
        var[cur] = sigsqr / (R(2.0) * omega) * (R(1.0) - rho * rho) + rho * rho * previous_var * (R(1.0) - var_ratio) +
                   yerr[cur] * yerr[cur];
        cur++;
    }
    void Predict(R tq, R* om, R* ov) {  // kfilter.cpp:72-135
        size_t ipredict = 0;
        while (tq > time[ipredict]) { ipredict++; if (ipredict == ny) break; }
        Reset();
        for (size_t i = 1; i < ipredict; i++) Update();
        R pm, pv, prec;
        if (ipredict == 0) { pm = 0; pv = sigsqr / (R(2.0) * omega); }
        else {
            R dtt = tq - time[ipredict - 1];
            R rho = std::exp(-dtt * omega);
            R previous_var = var[ipredict - 1] - yerr[ipredict - 1] * yerr[ipredict - 1];
            R var_ratio = previous_var / var[ipredict - 1];
            pm = rho * mean[ipredict - 1] + rho * var_ratio * (y[ipredict - 1] - mean[ipredict - 1]);
            pv = sigsqr / (R(2.0) * omega) * (R(1.0) - rho * rho) + rho * rho * previous_var * (R(1.0) - var_ratio);
        }
        if (ipredict == ny) { *om = pm; *ov = pv; return; }
        prec = R(1.0) / pv; pm *= prec;
        InitializeCoefs(tq, ipredict);
        prec += yslope * yslope / var[ipredict];
        pm += yslope * (y[ipredict] - yconst) / var[ipredict];
        for (size_t i = ipredict + 1; i < ny; i++) {
            UpdateCoefs();
            prec += yslope * yslope / var[i];
            pm += yslope * (y[i] - yconst) / var[i];
        }
        pv = R(1.0) / prec; pm *= pv;
        *om = pm; *ov = pv;
    }
};

// ---------------------------------------------------------------------------------
// CARMA_Base::LogDensity (carpack.hpp:131-176)
// ---------------------------------------------------------------------------------
template <class R>
struct Model {
    int kind = KIND_CARMA, p = 0, q = 0;
    Prior pr{};
    bool ignore_prior = false;
    std::vector<R> time, y, yerr;
    KFp<R> kf;
    KF1<R> kf1;

    int dim() const {
        if (kind == KIND_CAR1) return 4;
        if (kind == KIND_CARMA) return 3 + p + q;
        if (kind == KIND_ZCARMA) return 4 + p;
        return 3 + p;
    }

    void init(int k, int pp, int qq, const double* t, const double* yy, const double* ye, size_t n, const Prior& prior) {
        kind = k; p = (k == KIND_CAR1) ? 1 : pp; q = qq; pr = prior;
        time.resize(n); y.resize(n); yerr.resize(n);
        for (size_t i = 0; i < n; i++) { time[i] = (R)t[i]; y[i] = (R)yy[i]; yerr[i] = (R)ye[i]; }
    }

    R log_density(const R* theta) {
        const R ninf = -std::numeric_limits<R>::infinity();
        if (!check_prior_bounds<R>(kind, theta, p, pr, ignore_prior)) return ninf;
        R measerr_scale = theta[1], mu = theta[2];
        size_t n = time.size();
        std::vector<R> proposed_yerr(n), ycent(n);
        for (size_t i = 0; i < n; i++) { proposed_yerr[i] = std::sqrt(measerr_scale) * yerr[i]; ycent[i] = y[i] - mu; }
        const std::vector<R>* pmean; const std::vector<R>* pvar;
        if (kind == KIND_CAR1) {
            // carpack.hpp:265, 272-274
            R omega = std::exp(theta[3]);
            R sigsqr = R(2.0) * theta[0] * theta[0] * std::exp(theta[3]);
            kf1.sigsqr = sigsqr; kf1.omega = omega;
            kf1.set_series(time.data(), ycent.data(), proposed_yerr.data(), n);
            kf1.Filter();
            pmean = &kf1.mean; pvar = &kf1.var;
        } else {
            std::vector<cx<R>> omega;
            quad_roots<R>(theta + 3, p, omega);
            std::vector<R> ma = extract_ma<R>(kind, theta, p, q, pr);
            R sigsqr = theta[0] * theta[0] / carma_variance<R>(omega, ma, R(1.0), R(0.0));
            kf.set_series(time.data(), ycent.data(), proposed_yerr.data(), n);
            kf.set_params(sigsqr, omega, ma);
            if (!kf.Filter()) return ninf;  // arma::solve throw -> -inf (carpack.hpp:154-164)
            pmean = &kf.mean; pvar = &kf.var;
        }
        R logpost = 0;
        for (size_t i = 0; i < n; i++) {
            R yc = y[i] - (*pmean)[i] - mu;
            logpost += R(-0.5) * std::log((*pvar)[i]) - R(0.5) * yc * yc / (*pvar)[i];
        }
        logpost += log_prior<R>(kind, theta, p, pr);
        return logpost;
    }
};

// ---------------------------------------------------------------------------------
// Counter-based RNG shared (by definition, not by code) with the device path:
// Philox4x32-10.  The reference uses a global mt19937 (random.cpp:20); "seeded
// identically" is defined as identical (seed, stream, chain, iteration, slot)
// -> draw mapping (hard part 5).
// ---------------------------------------------------------------------------------
struct Philox {
    static void round_(uint32_t c[4], const uint32_t k[2]) {
        const uint64_t M0 = 0xD2511F53ull, M1 = 0xCD9E8D57ull;
        uint64_t p0 = M0 * c[0], p1 = M1 * c[2];
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    static void gen(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed, uint32_t out[4]) {
        uint32_t c[4] = {c0, c1, c2, c3};
        uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
        for (int r = 0; r < 10; r++) {
            round_(c, k);
            k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
        }
        for (int i = 0; i < 4; i++) out[i] = c[i];
    }
};

inline double u53(uint32_t a, uint32_t b) {
    // 53-bit uniform in (0,1): never 0 or 1
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}

enum Stream { STREAM_PROPOSAL = 0, STREAM_ACCEPT = 1, STREAM_EXCHANGE = 2, STREAM_START = 3 };

struct RngAddr {
    uint64_t seed; uint32_t chain;  // chain = global_ensemble * ntemps + temperature index
};

// two uniforms from block `blk` of (stream, iteration)
inline void uniforms2(const RngAddr& a, uint32_t stream, uint32_t iter, uint32_t blk, double* u0, double* u1) {
    uint32_t w[4];
    Philox::gen(blk, iter, a.chain, stream, a.seed, w);
    *u0 = u53(w[0], w[1]); *u1 = u53(w[2], w[3]);
}

inline double normal_from(double u0, double u1) {
    return std::sqrt(-2.0 * std::log(u0)) * std::cos(6.283185307179586476925286766559 * u1);
}

// Student-t with even dof:  N(0,1) / sqrt(chi2_dof / dof),  chi2_dof = -2 ln(prod_{dof/2} u)
// (same distribution as boost student_t_distribution used by random.cpp:158-164)
inline double tdist_draw(const RngAddr& a, uint32_t stream, uint32_t iter, uint32_t j, int dof) {
    int nblk = 1 + (dof / 2 + 1) / 2;
    uint32_t base = j * (uint32_t)nblk;
    double u0, u1;
    uniforms2(a, stream, iter, base, &u0, &u1);
    double z = normal_from(u0, u1);
    double prod = 1.0;
    int need = dof / 2;
    for (int b = 1; b < nblk; b++) {
        uniforms2(a, stream, iter, base + b, &u0, &u1);
        if (need > 0) { prod *= u0; need--; }
        if (need > 0) { prod *= u1; need--; }
    }
    double chi2 = -2.0 * std::log(prod);
    return z / std::sqrt(chi2 / (double)dof);
}

// chi-square with arbitrary integer dof for the starting values (random.cpp:180-186):
// even part via -2 ln prod u, odd remainder via one squared normal; long products are
// split in groups of 8 uniforms to stay inside double range.
struct StartRng {
    RngAddr a; uint32_t attempt; uint32_t blk = 0;
    void next2(double* u0, double* u1) { uniforms2(a, STREAM_START, attempt, blk++, u0, u1); }
    double uniform() { double u0, u1; next2(&u0, &u1); return u0; }
    double normal() { double u0, u1; next2(&u0, &u1); return normal_from(u0, u1); }
    double chisqr(int dof) {
        double acc = 0.0; int pairs = dof / 2; double prod = 1.0; int inprod = 0;
        for (int i = 0; i < pairs; i += 2) {
            double u0, u1; next2(&u0, &u1);
            prod *= u0; inprod++;
            if (i + 1 < pairs) { prod *= u1; inprod++; }
            if (inprod >= 8) { acc += -2.0 * std::log(prod); prod = 1.0; inprod = 0; }
        }
        if (inprod > 0) acc += -2.0 * std::log(prod);
        if (dof & 1) { double z = normal(); acc += z * z; }
        return acc;
    }
    double scaled_inverse_chisqr(int dof, double ssqr) { return ssqr / chisqr(dof) * (double)dof; }
};

// ---------------------------------------------------------------------------------
// Starting values (carpack.cpp:38-83 CAR1, 175-230 CARp, 268-311 StartingAR,
// 416-477 CARMA, 515-519 StartingMA, 586-644 ZCARMA, 681-684 StartingKappa)
// ---------------------------------------------------------------------------------
struct SeriesStats { double mean, var_sample, median_dt, min_dt, tmin, tmax; };

inline SeriesStats series_stats(const double* t, const double* y, size_t n) {
    SeriesStats s{};
    double sum = 0; for (size_t i = 0; i < n; i++) sum += y[i];
    s.mean = sum / (double)n;
    double ss = 0; for (size_t i = 0; i < n; i++) ss += (y[i] - s.mean) * (y[i] - s.mean);
    s.var_sample = ss / (double)(n - 1);  // arma::var default normalisation
    std::vector<double> dt(n - 1);
    for (size_t i = 0; i + 1 < n; i++) dt[i] = t[i + 1] - t[i];
    std::vector<double> sd(dt); std::sort(sd.begin(), sd.end());
    size_t m = sd.size();
    s.median_dt = (m % 2) ? sd[m / 2] : 0.5 * (sd[m / 2 - 1] + sd[m / 2]);
    s.min_dt = sd[0];
    s.tmin = t[0]; s.tmax = t[0];
    for (size_t i = 0; i < n; i++) { s.tmin = std::min(s.tmin, t[i]); s.tmax = std::max(s.tmax, t[i]); }
    return s;
}

inline void starting_ar(StartRng& g, int p, const Prior& pr, const SeriesStats& st, double* loga) {
    double min_freq = 1.0 / (st.tmax - st.tmin);
    int nl = (p + 1) / 2;
    std::vector<double> cent(nl), width(nl);
    for (int i = 0; i < nl; i++) cent[i] = std::exp(std::log(pr.max_freq / min_freq) * g.uniform() + std::log(min_freq));
    std::sort(cent.begin(), cent.end(), [](double a, double b) { return a > b; });
    for (int i = 0; i < nl; i++) width[i] = std::exp(std::log(pr.max_freq / min_freq) * g.uniform() + std::log(min_freq));
    if (p % 2 == 1) {
        cent[p / 2] = 0.0;
        double lo = std::log(min_freq);
        // carpack.cpp:289: uniform(log(min_freq), log(lorentz_cent(p/2-1))); for p==1 there is no such centroid
        double hi = (p / 2 >= 1) ? std::log(cent[p / 2 - 1]) : std::log(pr.max_freq);
        width[p / 2] = std::exp(lo + (hi - lo) * g.uniform());
    }
    const double pi = 3.14159265358979323846;
    for (int i = 0; i < p / 2; i++) {
        double real_part = -2.0 * pi * width[i];
        double imag_part = 2.0 * pi * cent[i];
        loga[2 * i] = std::log(real_part * real_part + imag_part * imag_part);
        loga[2 * i + 1] = std::log(-2.0 * real_part);
    }
    if (p % 2 == 1) loga[p - 1] = std::log(2.0 * pi * width[p / 2]);
}

// One attempt at a starting value.  Returns the log-posterior (may be non-finite).
inline double starting_value_attempt(Model<double>& m, const SeriesStats& st, StartRng& g, double* theta) {
    int p = m.p, q = m.q, kind = m.kind;
    size_t n = m.time.size();
    if (kind == KIND_CAR1) {
        double sd = std::sqrt(g.scaled_inverse_chisqr((int)n - 1, st.var_sample));
        double mu = st.mean + (sd / (double)n) * g.normal();
        double log_omega = -std::log(st.median_dt * (1.0 + 49.0 * g.uniform()));
        log_omega = std::min(log_omega, m.pr.max_freq);  // sic, carpack.cpp:56
        double scale = g.scaled_inverse_chisqr((int)m.pr.measerr_dof, 1.0);
        scale = std::max(std::min(scale, 1.99), 0.51);
        theta[0] = sd; theta[1] = scale; theta[2] = mu; theta[3] = log_omega;
        return m.log_density(theta);
    }
    starting_ar(g, p, m.pr, st, theta + 3);
    if (kind == KIND_CARMA) for (int i = 0; i < q; i++) theta[3 + p + i] = std::fabs(g.normal());
    if (kind == KIND_ZCARMA) theta[3 + p] = logit<double>(g.uniform());
    double yvar = g.scaled_inverse_chisqr((int)n - 1, st.var_sample);
    double mu = st.mean + (std::sqrt(yvar) / (double)n) * g.normal();
    double scale = g.scaled_inverse_chisqr((int)m.pr.measerr_dof, 1.0);
    scale = std::max(std::min(scale, 1.99), 0.51);
    theta[0] = std::sqrt(yvar); theta[1] = scale; theta[2] = mu;
    return m.log_density(theta);
}

// CholUpdateR1 (steps.cpp:111-131); L is the d x d upper factor, row-major.
inline void chol_update_r1(double* L, double* v, int d, bool downdate) {
    double sign = downdate ? -1.0 : 1.0;
    for (int k = 0; k < d; k++) {
        double lkk = L[k * d + k];
        double r = std::sqrt(lkk * lkk + sign * v[k] * v[k]);
        double c = r / lkk;
        double s = v[k] / lkk;
        L[k * d + k] = r;
        if (k < d - 1) {
            for (int j = k + 1; j < d; j++) L[k * d + j] = (L[k * d + j] + sign * s * v[j]) / c;
            for (int j = k + 1; j < d; j++) v[j] = c * v[j] - s * L[k * d + j];
        }
    }
}

}  // namespace

// ===================================================================================
// C entry points (ctypes)
// ===================================================================================
// "Lean" CPU variant of the same LogDensity: identical arithmetic in the identical order,
// but the filter state lives in fixed-size arrays (no heap traffic, no per-call vector copies, loops with
// compile-time bounds), mean/var are consumed on the fly instead of stored.  It bounds from above what the
// reference's algorithm can do on one core; the dense variant above is the one that mirrors its cost profile.
template <int P>
static double lean_filter_loglik(const double* dt, const double* y, const double* yerr, size_t ny, double sigsqr,
                                 const std::vector<cx<double>>& omega_v, const std::vector<double>& ma_v, double scale, double mu) {
    using C = cx<double>;
    C omega[P], b[P], x[P], K[P], rho[P], V[P][P], Pm[P][P];
    double ma[P];
    for (int i = 0; i < P; i++) { omega[i] = omega_v[i]; ma[i] = i < (int)ma_v.size() ? ma_v[i] : 0.0; x[i] = C(0, 0); }
    std::vector<C> E((size_t)P * P), J(P, C(0, 0));
    for (int k = 0; k < P; k++) E[k] = C(1, 0);
    if (P > 1) for (int k = 0; k < P; k++) E[(size_t)P + k] = omega[k];
    for (int i = 2; i < P; i++)
        for (int k = 0; k < P; k++) E[(size_t)i * P + k] = std::pow(omega[k], double(i));
    J[P - 1] = C(1, 0);
    if (!lu_solve<double>(E, P, J)) return -std::numeric_limits<double>::infinity();
    for (int k = 0; k < P; k++) {
        C s(0, 0);
        for (int i = 0; i < P; i++) s += ma[i] * E[(size_t)i * P + k];
        b[k] = s;
    }
    for (int i = 0; i < P; i++)
        for (int j = i; j < P; j++) V[i][j] = -sigsqr * J[i] * std::conj(J[j]) / (omega[i] + std::conj(omega[j]));
    for (int i = 0; i < P; i++)
        for (int j = 0; j < i; j++) V[i][j] = std::conj(V[j][i]);
    for (int i = 0; i < P; i++)
        for (int j = 0; j < P; j++) Pm[i][j] = V[i][j];
    auto quad = [&](C (*M)[P]) {
        C tot(0, 0);
        for (int j = 0; j < P; j++) {
            C bM(0, 0);
            for (int i = 0; i < P; i++) bM += b[i] * M[i][j];
            tot += bM * std::conj(b[j]);
        }
        return tot.real();
    };
    const double ssc = std::sqrt(scale);
    double mean = 0.0, e0 = ssc * yerr[0];
    double var = quad(V) + e0 * e0;
    double innovation = y[0] - mu;
    double ll = 0.0;
    for (size_t cur = 1;; cur++) {
        const double yc = y[cur - 1] - mean - mu;
        ll += -0.5 * std::log(var) - 0.5 * yc * yc / var;
        if (cur == ny) break;
        for (int i = 0; i < P; i++) {
            C s(0, 0);
            for (int j = 0; j < P; j++) s += Pm[i][j] * std::conj(b[j]);
            K[i] = s / var;
        }
        for (int i = 0; i < P; i++) x[i] += K[i] * innovation;
        for (int i = 0; i < P; i++)
            for (int j = 0; j < P; j++) Pm[i][j] -= var * (K[i] * std::conj(K[j]));
        for (int i = 0; i < P; i++) rho[i] = std::exp(omega[i] * dt[cur - 1]);
        for (int i = 0; i < P; i++) x[i] = rho[i] * x[i];
        for (int i = 0; i < P; i++)
            for (int j = 0; j < P; j++) Pm[i][j] = (rho[i] * std::conj(rho[j])) * (Pm[i][j] - V[i][j]) + V[i][j];
        C m(0, 0);
        for (int i = 0; i < P; i++) m += b[i] * x[i];
        mean = m.real();
        const double ec = ssc * yerr[cur];
        var = quad(Pm) + ec * ec;
        innovation = (y[cur] - mu) - mean;
    }
    return ll;
}

extern "C" {

struct oracle_prior { double max_stdev, max_freq, min_freq, kappa_low, kappa_high, measerr_dof; };

static Prior to_prior(const oracle_prior* p) {
    return Prior{p->max_stdev, p->max_freq, p->min_freq, p->kappa_low, p->kappa_high, p->measerr_dof};
}

// carpack.hpp:201-207 SetPrior + carpack.hpp:413-419 kappa bounds; max_stdev as in
// carmcmc.cpp:85-89 (population variance) when use_population_var != 0, else
// carpack.hpp:71 (arma::var, N-1).
void oracle_default_prior(const double* t, const double* y, size_t n, int use_population_var, oracle_prior* out) {
    SeriesStats st = series_stats(t, y, n);
    double var = st.var_sample;
    if (use_population_var) {
        double sum = 0, sq = 0;
        for (size_t i = 0; i < n; i++) { sum += y[i]; sq += y[i] * y[i]; }
        double mean = sum / (double)n;
        var = sq / (double)n - mean * mean;
    }
    out->max_stdev = 10.0 * std::sqrt(var);
    out->max_freq = 1.0 / st.min_dt;
    out->min_freq = 1.0 / (st.tmax - st.tmin);
    out->kappa_high = 1.0 / st.min_dt;
    out->kappa_low = std::max(1.0 / (st.tmax - st.tmin), 1.0 / (10.0 * st.median_dt));
    out->measerr_dof = 50.0;
}

void oracle_ar_roots(const double* logq, int p, double* roots_reim) {
    std::vector<cx<double>> r; quad_roots<double>(logq, p, r);
    for (int i = 0; i < p; i++) { roots_reim[2 * i] = r[i].real(); roots_reim[2 * i + 1] = r[i].imag(); }
}

void oracle_ma_coefs(int kind, const double* theta, int p, int q, const oracle_prior* pr, double* ma) {
    std::vector<double> m = extract_ma<double>(kind, theta, p, q, to_prior(pr));
    for (int i = 0; i < p; i++) ma[i] = m[i];
}

double oracle_variance(const double* roots_reim, const double* ma, int p, int nma, double sigma, double lag) {
    std::vector<cx<double>> r(p);
    for (int i = 0; i < p; i++) r[i] = cx<double>(roots_reim[2 * i], roots_reim[2 * i + 1]);
    std::vector<double> m(ma, ma + nma);
    return carma_variance<double>(r, m, sigma, lag);
}

int oracle_check_prior(int kind, const double* theta, int p, const oracle_prior* pr, int ignore_prior) {
    return check_prior_bounds<double>(kind, theta, p, to_prior(pr), ignore_prior != 0) ? 1 : 0;
}

double oracle_log_prior(int kind, const double* theta, int p, const oracle_prior* pr) {
    return log_prior<double>(kind, theta, p, to_prior(pr));
}

// KalmanFilterp::Filter with explicit (sigsqr, omega, ma): kfilter.hpp:303-334 + Filter().
// y must already be centred, yerr are standard deviations.  Returns 0, or 1 on a singular solve.
int oracle_filterp(const double* t, const double* y, const double* yerr, size_t ny, double sigsqr,
                   const double* omega_reim, const double* ma, int p, double* mean, double* var) {
    KFp<double> kf;
    kf.set_series(t, y, yerr, ny);
    std::vector<cx<double>> om(p);
    for (int i = 0; i < p; i++) om[i] = cx<double>(omega_reim[2 * i], omega_reim[2 * i + 1]);
    kf.set_params(sigsqr, om, std::vector<double>(ma, ma + p));
    if (!kf.Filter()) return 1;
    for (size_t i = 0; i < ny; i++) { mean[i] = kf.mean[i]; var[i] = kf.var[i]; }
    return 0;
}

int oracle_predictp(const double* t, const double* y, const double* yerr, size_t ny, double sigsqr,
                    const double* omega_reim, const double* ma, int p, const double* tq, size_t nq,
                    double* qmean, double* qvar) {
    KFp<double> kf;
    kf.set_series(t, y, yerr, ny);
    std::vector<cx<double>> om(p);
    for (int i = 0; i < p; i++) om[i] = cx<double>(omega_reim[2 * i], omega_reim[2 * i + 1]);
    kf.set_params(sigsqr, om, std::vector<double>(ma, ma + p));
    for (size_t k = 0; k < nq; k++)
        if (!kf.Predict(tq[k], &qmean[k], &qvar[k])) return 1;
    return 0;
}

void oracle_filter1(const double* t, const double* y, const double* yerr, size_t ny, double sigsqr, double omega,
                    double* mean, double* var) {
    KF1<double> kf; kf.sigsqr = sigsqr; kf.omega = omega;
    kf.set_series(t, y, yerr, ny); kf.Filter();
    for (size_t i = 0; i < ny; i++) { mean[i] = kf.mean[i]; var[i] = kf.var[i]; }
}

void oracle_predict1(const double* t, const double* y, const double* yerr, size_t ny, double sigsqr, double omega,
                     const double* tq, size_t nq, double* qmean, double* qvar) {
    KF1<double> kf; kf.sigsqr = sigsqr; kf.omega = omega;
    kf.set_series(t, y, yerr, ny);
    for (size_t k = 0; k < nq; k++) kf.Predict(tq[k], &qmean[k], &qvar[k]);
}

// CARMA_Base::LogDensity for n parameter vectors (row-major n x d), double precision.
void oracle_logdensity_batch(int kind, int p, int q, const double* t, const double* y, const double* yerr, size_t ny,
                             const oracle_prior* pr, int ignore_prior, const double* theta, size_t n, double* out) {
    Model<double> m;
    m.init(kind, p, q, t, y, yerr, ny, to_prior(pr));
    m.ignore_prior = ignore_prior != 0;
    int d = m.dim();
    for (size_t i = 0; i < n; i++) out[i] = m.log_density(theta + i * (size_t)d);
}

// Same in long double; used only to measure the intrinsic rounding noise of the algorithm.
void oracle_logdensity_batch_ld(int kind, int p, int q, const double* t, const double* y, const double* yerr, size_t ny,
                                const oracle_prior* pr, int ignore_prior, const double* theta, size_t n, double* out) {
    Model<long double> m;
    m.init(kind, p, q, t, y, yerr, ny, to_prior(pr));
    m.ignore_prior = ignore_prior != 0;
    int d = m.dim();
    std::vector<long double> th(d);
    for (size_t i = 0; i < n; i++) {
        for (int j = 0; j < d; j++) th[j] = (long double)theta[i * (size_t)d + j];
        out[i] = (double)m.log_density(th.data());
    }
}

// CARMA kinds with p >= 2 only (the benchmark shape); same bounds / theta transform / prior code as the dense variant.
void oracle_logdensity_batch_lean(int kind, int p, int q, const double* t, const double* y, const double* yerr, size_t ny,
                                  const oracle_prior* pr_, int ignore_prior, const double* theta, size_t n, double* out) {
    const Prior pr = to_prior(pr_);
    std::vector<double> dt(ny > 1 ? ny - 1 : 0);
    for (size_t i = 0; i + 1 < ny; i++) dt[i] = t[i + 1] - t[i];
    const int d = (kind == KIND_CARMA) ? 3 + p + q : (kind == KIND_ZCARMA ? 4 + p : 3 + p);
    std::vector<cx<double>> omega;
    for (size_t r = 0; r < n; r++) {
        const double* th = theta + r * (size_t)d;
        if (p < 2 || p > 7 || !check_prior_bounds<double>(kind, th, p, pr, ignore_prior != 0)) {
            out[r] = -std::numeric_limits<double>::infinity();
            continue;
        }
        quad_roots<double>(th + 3, p, omega);
        std::vector<double> ma = extract_ma<double>(kind, th, p, q, pr);
        const double sigsqr = th[0] * th[0] / carma_variance<double>(omega, ma, 1.0, 0.0);
        double ll;
        switch (p) {
            case 2: ll = lean_filter_loglik<2>(dt.data(), y, yerr, ny, sigsqr, omega, ma, th[1], th[2]); break;
            case 3: ll = lean_filter_loglik<3>(dt.data(), y, yerr, ny, sigsqr, omega, ma, th[1], th[2]); break;
            case 4: ll = lean_filter_loglik<4>(dt.data(), y, yerr, ny, sigsqr, omega, ma, th[1], th[2]); break;
            case 5: ll = lean_filter_loglik<5>(dt.data(), y, yerr, ny, sigsqr, omega, ma, th[1], th[2]); break;
            case 6: ll = lean_filter_loglik<6>(dt.data(), y, yerr, ny, sigsqr, omega, ma, th[1], th[2]); break;
            default: ll = lean_filter_loglik<7>(dt.data(), y, yerr, ny, sigsqr, omega, ma, th[1], th[2]); break;
        }
        out[r] = ll + log_prior<double>(kind, th, p, pr);
    }
}

// Multi-series batch: curve c owns rows [off[c], off[c+1]) of t/y/yerr and one theta row.
void oracle_logdensity_multi(int kind, int p, int q, const double* t, const double* y, const double* yerr,
                             const int64_t* off, size_t ncurves, const oracle_prior* priors, int ignore_prior,
                             const double* theta, double* out) {
    for (size_t c = 0; c < ncurves; c++) {
        Model<double> m;
        size_t n = (size_t)(off[c + 1] - off[c]);
        m.init(kind, p, q, t + off[c], y + off[c], yerr + off[c], n, to_prior(&priors[c]));
        m.ignore_prior = ignore_prior != 0;
        out[c] = m.log_density(theta + c * (size_t)m.dim());
    }
}

// raw Philox block, and derived draws, so that tests can pin the device RNG bit for bit
void oracle_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed, uint32_t* out4) {
    Philox::gen(c0, c1, c2, c3, seed, out4);
}
double oracle_tdist(uint64_t seed, uint32_t chain, uint32_t iter, uint32_t j, int dof) {
    RngAddr a{seed, chain};
    return tdist_draw(a, STREAM_PROPOSAL, iter, j, dof);
}

void oracle_chol_update(double* L, double* v, int d, int downdate) { chol_update_r1(L, v, d, downdate != 0); }

// Starting value for (seed, chain): redraw until finite (carpack.cpp:182-227), capped.
int oracle_starting_value(int kind, int p, int q, const double* t, const double* y, const double* yerr, size_t ny,
                          const oracle_prior* pr, uint64_t seed, uint32_t chain, int max_attempts, double* theta,
                          double* logpost) {
    Model<double> m;
    m.init(kind, p, q, t, y, yerr, ny, to_prior(pr));
    SeriesStats st = series_stats(t, y, ny);
    for (int a = 0; a < max_attempts; a++) {
        StartRng g{RngAddr{seed, chain}, (uint32_t)a};
        double lp = starting_value_attempt(m, st, g, theta);
        if (std::isfinite(lp)) { *logpost = lp; return a; }
    }
    *logpost = -std::numeric_limits<double>::infinity();
    return -1;
}

struct oracle_pt_opts {
    int nsamples, burnin, thin, ntemps;
    double tmax;        // carmcmc.cpp:92 (100)
    int dof;            // carmcmc.cpp:139 (8)
    double target_rate; // carmcmc.cpp:141 (0.25)
    double gamma;       // steps.cpp:29 (2/3)
    uint64_t seed;
    uint32_t ensemble;  // global ensemble index (chain id = ensemble*ntemps + temperature)
    int max_start_attempts;
};

// One record per RAM step / exchange step, for record-replay parity tests.
struct oracle_trace_rec {
    double lp_prop;   // LogDensity(proposal)
    double lp_cur;    // cached log-posterior of the chain before the step
    double alpha;     // acceptance probability after Accept() (0 when non-finite)
    double u;         // uniform used (NaN if not drawn)
    int accepted;
    int pad;
};

// RunCarmaSampler / RunCar1Sampler (carmcmc.cpp:30-177) + Sampler::Run (samplers.cpp:57-115)
// + AdaptiveMetro (steps.cpp:24-107) + ExchangeStep (steps.hpp:318-362), driven by Philox.
// init: d values or NULL.  samples: nsamples x d, logposts: nsamples (coolest chain).
// trace (optional): (burnin + nsamples*thin) x ntemps RAM records followed by the same count
// of exchange records (index [iter][i], exchange i<->i-1 stored at i; i==0 unused).
// theta_trace (optional): proposals, (iters x ntemps x d).
int oracle_pt_run(int kind, int p, int q, const double* t, const double* y, const double* yerr, size_t ny,
                  const oracle_prior* pr, const oracle_pt_opts* o, const double* init, double* samples,
                  double* logposts, double* accept_rates, oracle_trace_rec* trace, double* theta_trace,
                  double* final_chol) {
    int T = o->ntemps;
    std::vector<Model<double>> models(T);
    for (int i = 0; i < T; i++) models[i].init(kind, p, q, t, y, yerr, ny, to_prior(pr));
    int d = models[0].dim();
    SeriesStats st = series_stats(t, y, ny);
    // carmcmc.cpp:85-89 population variance of y
    double sum = 0, sq = 0;
    for (size_t i = 0; i < ny; i++) { sum += y[i]; sq += y[i] * y[i]; }
    double mean = sum / (double)ny;
    double var = sq / (double)ny - mean * mean;
    // carmcmc.cpp:92-95 ladder
    std::vector<double> temp(T, 1.0);
    for (int i = 0; i < T; i++) temp[i] = (T > 1) ? std::exp(std::log(o->tmax) * (double)i / (double)(T - 1)) : 1.0;
    // carmcmc.cpp:127-136 initial proposal covariance (diagonal) -> upper Cholesky factor
    std::vector<std::vector<double>> R(T, std::vector<double>((size_t)d * d, 0.0));
    for (int c = 0; c < T; c++) {
        for (int j = 0; j < d; j++) R[c][(size_t)j * d + j] = 0.01;
        R[c][0] = std::sqrt(2.0 * var * var / (double)ny);
        R[c][(size_t)2 * d + 2] = std::sqrt(var / (double)ny);
    }
    std::vector<std::vector<double>> theta(T, std::vector<double>(d));
    std::vector<double> lp(T);
    std::vector<int> niter(T, 0), naccept(T, 0);
    // samplers.cpp:75-93 starting values (once per chain)
    for (int c = 0; c < T; c++) {
        bool ok = false;
        if (init) {
            double l = models[c].log_density(init);
            if (std::isfinite(l)) { for (int j = 0; j < d; j++) theta[c][j] = init[j]; lp[c] = l; ok = true; }
        }
        if (!ok) {
            int a;
            for (a = 0; a < o->max_start_attempts; a++) {
                StartRng g{RngAddr{o->seed, o->ensemble * (uint32_t)T + (uint32_t)c}, (uint32_t)a};
                double l = starting_value_attempt(models[c], st, g, theta[c].data());
                if (std::isfinite(l)) { lp[c] = l; break; }
            }
            if (a == o->max_start_attempts) return 2;
        }
    }
    int total_iters = o->burnin + o->nsamples * o->thin;
    std::vector<double> z(d), sp(d), nv(d);
    int nsaved = 0;
    for (int it = 0; it < total_iters; it++) {
        for (int c = T - 1; c >= 0; c--) {
            RngAddr addr{o->seed, o->ensemble * (uint32_t)T + (uint32_t)c};
            // ---- AdaptiveMetro::DoStep (steps.cpp:60-107)
            for (int j = 0; j < d; j++) z[j] = tdist_draw(addr, STREAM_PROPOSAL, (uint32_t)it, (uint32_t)j, o->dof);
            for (int j = 0; j < d; j++) {  // chol_factor_.t() * unit_proposal
                double s = 0;
                for (int k = 0; k <= j; k++) s += R[c][(size_t)k * d + j] * z[k];
                sp[j] = s; nv[j] = theta[c][j] + s;
            }
            double lpn = models[c].log_density(nv.data());
            double alpha = (lpn - lp[c]) / temp[c];
            double u = std::numeric_limits<double>::quiet_NaN();
            bool acc = false;
            if (!std::isfinite(alpha)) { alpha = 0.0; }
            else {
                double u0, u1; uniforms2(addr, STREAM_ACCEPT, (uint32_t)it, 0, &u0, &u1); u = u0;
                alpha = std::min(std::exp(alpha), 1.0);
                if (u < alpha) { naccept[c]++; acc = true; }
            }
            if (trace) {
                oracle_trace_rec& r = trace[(size_t)it * T + c];
                r.lp_prop = lpn; r.lp_cur = lp[c]; r.alpha = alpha; r.u = u; r.accepted = acc; r.pad = 0;
            }
            if (theta_trace) for (int j = 0; j < d; j++) theta_trace[((size_t)it * T + c) * d + j] = nv[j];
            if (acc) { theta[c] = nv; lp[c] = lpn; }
            if (niter[c] < o->burnin) {
                double step = std::min(1.0, (double)d / std::pow((double)niter[c], o->gamma));
                double nrm = 0; for (int j = 0; j < d; j++) nrm += z[j] * z[j]; nrm = std::sqrt(nrm);
                double f = std::sqrt(step * std::fabs(alpha - o->target_rate)) / nrm;
                for (int j = 0; j < d; j++) sp[j] = f * sp[j];
                chol_update_r1(R[c].data(), sp.data(), d, alpha < o->target_rate);
            }
            niter[c]++;
            // ---- ExchangeStep::DoStep (steps.hpp:318-362): chain c <-> c-1
            if (c > 0) {
                double this_lp = lp[c], other_lp = lp[c - 1];
                double a = 1.0 / temp[c] * (other_lp - this_lp) + 1.0 / temp[c - 1] * (this_lp - other_lp);
                double u0, u1; uniforms2(addr, STREAM_EXCHANGE, (uint32_t)it, 0, &u0, &u1);
                a = std::min(std::exp(a), 1.0);
                if (!std::isfinite(a)) a = 0.0;
                bool sw = u0 < a;
                if (sw) { std::swap(theta[c], theta[c - 1]); std::swap(lp[c], lp[c - 1]); }
                if (trace) {
                    oracle_trace_rec& r = trace[(size_t)total_iters * T + (size_t)it * T + c];
                    r.lp_prop = other_lp; r.lp_cur = this_lp; r.alpha = a; r.u = u0; r.accepted = sw; r.pad = 0;
                }
            }
        }
        // samplers.cpp:101-108: after each `thin` iterations past burn-in store chain 0
        if (it >= o->burnin && ((it - o->burnin + 1) % o->thin) == 0 && nsaved < o->nsamples) {
            for (int j = 0; j < d; j++) samples[(size_t)nsaved * d + j] = theta[0][j];
            logposts[nsaved] = lp[0];
            nsaved++;
        }
    }
    if (accept_rates) for (int c = 0; c < T; c++) accept_rates[c] = niter[c] ? (double)naccept[c] / niter[c] : 0.0;
    if (final_chol) for (int c = 0; c < T; c++) std::memcpy(final_chol + (size_t)c * d * d, R[c].data(), sizeof(double) * d * d);
    return 0;
}

}  // extern "C"
