// kalman_cplx.cuh -- general complex-Hermitian CARMA Kalman filter with explicit (sigsqr, omega, ma),
// for the KalmanFilterp / KalmanFilter1 class API (Filter, GetMean/GetVar, Predict).  Unlike the
// log-density kernels it makes no assumption on the roots (any order, any complex values), so it
// keeps the full rotated complex state exactly as the reference writes it.
//
// Reference (paths relative to /root/reference/src):
//   KalmanFilterp::Reset            kfilter.cpp:138-186
//   KalmanFilterp::Update           kfilter.cpp:189-215
//   KalmanFilterp::Predict          kfilter.cpp:218-286
//   KalmanFilterp::InitializeCoefs  kfilter.cpp:290-312
//   KalmanFilterp::UpdateCoefs      kfilter.cpp:316-337
#pragma once
#include "theta_transform.cuh"

namespace carma {

struct SeriesView {
    const double* dt;   // dt[i] = t[i+1]-t[i], dt[ny-1] = 0
    const double* y;    // y[i]
    const double* e2n;  // yerr[i+1]^2, e2n[ny-1] = 0
    const double* t;    // t[i]
    double e2_0;        // yerr[0]^2
    double dt_max;      // longest gap (rate clamp of transform_theta)
    int ny;
    int nyp;            // padded length (even) of dt / y / e2n
};

struct KalmanCplx {
    int p;
    cxd w[MAX_P], b[MAX_P], x[MAX_P], g[MAX_P];
    cxd V[MAX_P][MAX_P], Pm[MAX_P][MAX_P];
    cxd sc[MAX_P], ss[MAX_P];  // state_const_, state_slope_
    double yconst, yslope;
    double var, mean, innov;
    double scale, mu;

    // kfilter.cpp:138-186; returns false when the Vandermonde system is singular
    __device__ bool reset(double sigsqr, const double* omega_reim, const double* ma, int p_, double e2_0, double y0) {
        p = p_;
        for (int k = 0; k < p; k++) w[k] = cx(omega_reim[2 * k], omega_reim[2 * k + 1]);
        for (int k = 0; k < p; k++) {
            cxd s = cx(ma[p - 1], 0);
            for (int l = p - 2; l >= 0; l--) s = s * w[k] + cx(ma[l], 0);
            b[k] = s;
            x[k] = cx(0, 0);
        }
        // J = E^{-1} e_p by LU with partial pivoting (arma::solve -> zgesv, kfilter.cpp:152-158)
        cxd J[MAX_P];
        {
            cxd A[MAX_P][MAX_P];
            for (int k = 0; k < p; k++) {
                cxd pw = cx(1, 0);
                A[0][k] = pw;
                for (int i = 1; i < p; i++) { pw = pw * w[k]; A[i][k] = pw; }
                J[k] = cx(k == p - 1 ? 1.0 : 0.0, 0.0);
            }
            for (int k = 0; k < p; k++) {
                int piv = k;
                double best = fabs(A[k][k].re) + fabs(A[k][k].im);
                for (int i = k + 1; i < p; i++) {
                    double v = fabs(A[i][k].re) + fabs(A[i][k].im);
                    if (v > best) { best = v; piv = i; }
                }
                if (!(best > 0.0) || !isfinite(best)) return false;
                if (piv != k) {
                    for (int j = 0; j < p; j++) { cxd tmp = A[k][j]; A[k][j] = A[piv][j]; A[piv][j] = tmp; }
                    cxd tr = J[k]; J[k] = J[piv]; J[piv] = tr;
                }
                for (int i = k + 1; i < p; i++) {
                    cxd l = cdiv(A[i][k], A[k][k]);
                    for (int j = k + 1; j < p; j++) A[i][j] = A[i][j] - l * A[k][j];
                    J[i] = J[i] - l * J[k];
                }
            }
            for (int i = p - 1; i >= 0; i--) {
                cxd s = J[i];
                for (int j = i + 1; j < p; j++) s = s - A[i][j] * J[j];
                J[i] = cdiv(s, A[i][i]);
            }
        }
        for (int i = 0; i < p; i++)
            for (int j = i; j < p; j++) {
                cxd v = cdiv((-sigsqr) * (J[i] * conj(J[j])), w[i] + conj(w[j]));
                V[i][j] = v;
                V[j][i] = conj(v);
            }
        for (int i = 0; i < p; i++)
            for (int j = 0; j < p; j++) Pm[i][j] = V[i][j];
        mean = 0.0;
        var = quad_form() + scale * e2_0;
        innov = (y0 - mu);
        return true;
    }

    // g = P b^H and Re(b g)
    __device__ double quad_form() {
        double tot = 0.0;
        for (int i = 0; i < p; i++) {
            cxd s = cx(0, 0);
            for (int j = 0; j < p; j++) s = s + Pm[i][j] * conj(b[j]);
            g[i] = s;
            tot += b[i].re * s.re - b[i].im * s.im;
        }
        return tot;
    }

    // kfilter.cpp:191-204 (g must hold P b^H): gain, state update, covariance update, transition by dt
    __device__ void gain_and_advance(double var_prev, double dtt) {
        double inv = 1.0 / var_prev;
        for (int i = 0; i < p; i++) x[i] = x[i] + (innov * inv) * g[i];
        for (int i = 0; i < p; i++)
            for (int j = 0; j < p; j++) Pm[i][j] = Pm[i][j] - inv * (g[i] * conj(g[j]));
        cxd rho[MAX_P];
        for (int i = 0; i < p; i++) {
            double e = exp(w[i].re * dtt), sn, cs;
            sincos(w[i].im * dtt, &sn, &cs);
            rho[i] = cx(e * cs, e * sn);
            x[i] = rho[i] * x[i];
        }
        for (int i = 0; i < p; i++)
            for (int j = 0; j < p; j++) Pm[i][j] = (rho[i] * conj(rho[j])) * (Pm[i][j] - V[i][j]) + V[i][j];
    }

    // kfilter.cpp:189-215
    __device__ void update(double dtt, double y_next, double e2_next) {
        gain_and_advance(var, dtt);
        double m = 0.0;
        for (int i = 0; i < p; i++) m += b[i].re * x[i].re - b[i].im * x[i].im;
        mean = m;
        var = quad_form() + scale * e2_next;
        innov = (y_next - mu) - mean;
    }

    // kfilter.cpp:290-312  (g must hold P b^H)
    __device__ void initialize_coefs(double dtt, double ymean, double yvar, double e2_at) {
        double inv = 1.0 / yvar;
        for (int i = 0; i < p; i++) {
            cxd K = inv * g[i];
            sc[i] = x[i] - ymean * K;
            ss[i] = K;
        }
        for (int i = 0; i < p; i++)
            for (int j = 0; j < p; j++) Pm[i][j] = Pm[i][j] - inv * (g[i] * conj(g[j]));
        coef_transition(dtt, e2_at);
    }

    __device__ void coef_transition(double dtt, double e2_at) {
        cxd rho[MAX_P];
        for (int i = 0; i < p; i++) {
            double e = exp(w[i].re * dtt), sn, cs;
            sincos(w[i].im * dtt, &sn, &cs);
            rho[i] = cx(e * cs, e * sn);
            sc[i] = rho[i] * sc[i];
            ss[i] = rho[i] * ss[i];
        }
        for (int i = 0; i < p; i++)
            for (int j = 0; j < p; j++) Pm[i][j] = (rho[i] * conj(rho[j])) * (Pm[i][j] - V[i][j]) + V[i][j];
        double c = 0.0, s = 0.0;
        for (int i = 0; i < p; i++) {
            c += b[i].re * sc[i].re - b[i].im * sc[i].im;
            s += b[i].re * ss[i].re - b[i].im * ss[i].im;
        }
        yconst = c;
        yslope = s;
        var = quad_form() + scale * e2_at;
    }

    // kfilter.cpp:316-337
    __device__ void update_coefs(double dtt, double y_prev, double e2_at) {
        double inv = 1.0 / var;
        for (int i = 0; i < p; i++) {
            cxd K = inv * g[i];
            sc[i] = sc[i] + ((y_prev - mu) - yconst) * K;
            ss[i] = ss[i] - yslope * K;
        }
        for (int i = 0; i < p; i++)
            for (int j = 0; j < p; j++) Pm[i][j] = Pm[i][j] - inv * (g[i] * conj(g[j]));
        coef_transition(dtt, e2_at);
    }
};

}  // namespace carma
