// kalman_real.cuh -- the CARMA(p,q) Kalman filter recursion, one filter per thread, whole state
// in registers (K1 in DESIGN.md).
//
// Reference recursion restated (paths relative to /root/reference/src):
//   KalmanFilterp::Reset   kfilter.cpp:138-186     P = V, x = 0, var_0 = Re(b V b^H) + yerr_0^2
//   KalmanFilterp::Update  kfilter.cpp:189-215     K = P b^H / var;  x += K innov;  P -= var K K^H;
//                                                  rho = exp(omega dt); x = rho o x;
//                                                  P = (rho rho^H) o (P - V) + V;
//                                                  mean = Re(b x); var = Re(b P b^H) + yerr^2
//   CARMA_Base::LogDensity carpack.hpp:167-171     ll = sum -1/2 log var_i - 1/2 (y_i - mean_i - mu)^2 / var_i
//
// What is different from a transliteration (all exact algebraic identities):
//   * the state is the independent real half z of the rotated complex state (theta_transform.cuh);
//     the Hadamard product (rho rho^H) o . becomes the block-diagonal congruence Phi . Phi^T with 2x2
//     blocks  e^{a dt} [[cos, -sin],[sin, cos]]  (conjugate pair) or diag(e^{r1 dt}, e^{r2 dt}) (real pair);
//   * D = P - V is stored instead of P, so V never enters the time loop: g = P b^H = D c + h,
//     var = c.g + yerr^2, and the predict step is D <- Phi D Phi^T (no subtract/add of V);
//   * g is computed once per step and reused for the gain, the state update and the variance.
#pragma once
#include "fast_math.cuh"
#include "theta_transform.cuh"

namespace carma {

template <int P>
struct KalmanReal {
    static constexpr int NS = P / 2;         // 2x2 slots
    static constexpr bool ODD = (P & 1) != 0;
    static constexpr int NT = P * (P + 1) / 2;

    double D[NT];  // upper triangle of the symmetric D = P - V
    double z[P];
    double g[P];   // P b^H in the real basis (before division by var)
    double var, mean;

    static __host__ __device__ constexpr int idx(int i, int j) { return i * P - (i * (i - 1)) / 2 + (j - i); }

    __device__ __forceinline__ void reset(const RealParams<P>& prm, double e2_0) {
#pragma unroll
        for (int i = 0; i < NT; i++) D[i] = 0.0;
#pragma unroll
        for (int i = 0; i < P; i++) { z[i] = 0.0; g[i] = prm.h[i]; }
        mean = 0.0;
        var = prm.v0 + prm.scale * e2_0;
    }

    // One Update(): condition on the residual `innov` (= y_i - mu - mean_i) observed with predictive
    // variance `var`, move forward by dt, and form mean/var for the next point (measurement variance e2n).
    // ALLC = true: every 2x2 slot is a conjugate pair (compile-time straight-line code, the common case);
    // ALLC = false: per-slot run-time selection between conjugate pair and real pair.
    template <bool ALLC>
    __device__ __forceinline__ void advance(const RealParams<P>& prm, double innov, double inv_var, double dt,
                                            double e2n) {
        measurement_update(innov, inv_var);
        predict_observe<ALLC>(prm, dt, e2n);
    }

    // z += g innov/var ;  D -= g g^T / var   (kfilter.cpp:191-197)
    __device__ __forceinline__ void measurement_update(double innov, double inv_var) {
        const double w = innov * inv_var;
        double gi[P];
#pragma unroll
        for (int i = 0; i < P; i++) {
            z[i] = fma(g[i], w, z[i]);
            gi[i] = g[i] * inv_var;
        }
#pragma unroll
        for (int i = 0; i < P; i++)
#pragma unroll
            for (int j = i; j < P; j++) D[idx(i, j)] = fma(-gi[i], g[j], D[idx(i, j)]);
    }

    // transition by dt and predicted observation of the next point (kfilter.cpp:200-210)
    template <bool ALLC>
    __device__ __forceinline__ void predict_observe(const RealParams<P>& prm, double dt, double e2n) {
        // ---- transition blocks Phi_s
        double f00[NS > 0 ? NS : 1], f01[NS > 0 ? NS : 1], f10[NS > 0 ? NS : 1], f11[NS > 0 ? NS : 1];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            // the decay factor of the first root is common to both slot types
            double e = exp_fast(prm.lam[2 * s] * dt);
            double sn, cs;
            sincos_fast(prm.lam[2 * s + 1] * dt, &sn, &cs);
            if (ALLC) {
                f00[s] = e * cs; f01[s] = -(e * sn); f10[s] = e * sn; f11[s] = e * cs;
            } else {
                // generic loop: BOTH candidates are evaluated for every lane (lam[2s+1] <= 0 in either
                // reading, so the extra exp is harmless) and selected per lane -- straight-line code.  A
                // per-lane branch here costs more than the 12 extra FP64 instructions: mixed warps would
                // run both sides anyway and the split basic blocks stop ptxas from interleaving the chains
                // (measured: PT-MCMC config 3, 159 -> 140 ms).
                const bool is_c = (prm.cmask >> s) & 1u;
                const double e2 = exp_fast(prm.lam[2 * s + 1] * dt);
                const double ec = e * cs, es = e * sn;
                f00[s] = is_c ? ec : e;
                f01[s] = is_c ? -es : 0.0;
                f10[s] = is_c ? es : 0.0;
                f11[s] = is_c ? ec : e2;
            }
        }
        double fo = 1.0;
        if (ODD) fo = exp_fast(prm.lam[P - 1] * dt);

        // ---- predict state
#pragma unroll
        for (int s = 0; s < NS; s++) {
            double u = z[2 * s], v = z[2 * s + 1];
            z[2 * s] = fma(f00[s], u, f01[s] * v);
            z[2 * s + 1] = fma(f10[s], u, f11[s] * v);
        }
        if (ODD) z[P - 1] *= fo;

        // ---- predict covariance: D <- Phi D Phi^T, block by block
#pragma unroll
        for (int a = 0; a < NS; a++) {
            const int i = 2 * a;
            {   // diagonal block (symmetric 2x2)
                double d00 = D[idx(i, i)], d01 = D[idx(i, i + 1)], d11 = D[idx(i + 1, i + 1)];
                double m00 = fma(f00[a], d00, f01[a] * d01), m01 = fma(f00[a], d01, f01[a] * d11);
                double m10 = fma(f10[a], d00, f11[a] * d01), m11 = fma(f10[a], d01, f11[a] * d11);
                D[idx(i, i)] = fma(m00, f00[a], m01 * f01[a]);
                D[idx(i, i + 1)] = fma(m00, f10[a], m01 * f11[a]);
                D[idx(i + 1, i + 1)] = fma(m10, f10[a], m11 * f11[a]);
            }
#pragma unroll
            for (int b = a + 1; b < NS; b++) {  // off-diagonal 2x2 block
                const int j = 2 * b;
                double d00 = D[idx(i, j)], d01 = D[idx(i, j + 1)], d10 = D[idx(i + 1, j)], d11 = D[idx(i + 1, j + 1)];
                double m00 = fma(f00[a], d00, f01[a] * d10), m01 = fma(f00[a], d01, f01[a] * d11);
                double m10 = fma(f10[a], d00, f11[a] * d10), m11 = fma(f10[a], d01, f11[a] * d11);
                D[idx(i, j)] = fma(m00, f00[b], m01 * f01[b]);
                D[idx(i, j + 1)] = fma(m00, f10[b], m01 * f11[b]);
                D[idx(i + 1, j)] = fma(m10, f00[b], m11 * f01[b]);
                D[idx(i + 1, j + 1)] = fma(m10, f10[b], m11 * f11[b]);
            }
            if (ODD) {  // 2x1 block against the odd real root
                double d0 = D[idx(i, P - 1)] * fo, d1 = D[idx(i + 1, P - 1)] * fo;
                D[idx(i, P - 1)] = fma(f00[a], d0, f01[a] * d1);
                D[idx(i + 1, P - 1)] = fma(f10[a], d0, f11[a] * d1);
            }
        }
        if (ODD) D[idx(P - 1, P - 1)] *= fo * fo;

        // ---- predicted observation: g = D c + h, var = c.g + e2, mean = c.z
        // __dmul_rn: never contracted into the following add, so the all-conjugate-pairs loop and the
        // generic loop round identically (results must not depend on which lanes share a warp)
        double m = 0.0, vv = __dmul_rn(prm.scale, e2n);
        if (ALLC) {
            // c = (1,0, 1,0, ..., [1]): plain sums over the first component of every slot
#pragma unroll
            for (int i = 0; i < P; i++) {
                double acc = prm.h[i];
#pragma unroll
                for (int j = 0; j < P; j++)
                    if ((j & 1) == 0) acc += D[(i <= j) ? idx(i, j) : idx(j, i)];
                g[i] = acc;
            }
#pragma unroll
            for (int i = 0; i < P; i++)
                if ((i & 1) == 0) { vv += g[i]; m += z[i]; }
        } else {
#pragma unroll
            for (int i = 0; i < P; i++) {
                double acc = prm.h[i];
#pragma unroll
                for (int j = 0; j < P; j++) acc = fma(D[(i <= j) ? idx(i, j) : idx(j, i)], prm.c[j], acc);
                g[i] = acc;
                vv = fma(prm.c[i], acc, vv);
                m = fma(prm.c[i], z[i], m);
            }
        }
        var = vv;
        mean = m;
    }
};

// Running sum of -1/2 log(var_i) - 1/2 innov_i^2 / var_i with the logs folded into one log of a
// product of mantissas (exponents summed as integers): log() leaves the time loop.
// out-of-line: keeps the (never taken in practice) log() code out of the time loop's instruction footprint
static __device__ __noinline__ double loglik_slow_log(double var) { return log(var); }

struct LogLikAcc {
    double quad;    // sum innov^2 / var
    double prod;    // product of mantissas of var, renormalised
    double logsum;  // direct sum of log(var) for out-of-range var (rare)
    int esum;
    int n_in_prod;
    __device__ __forceinline__ void init() { quad = 0.0; prod = 1.0; logsum = 0.0; esum = 0; n_in_prod = 0; }
    __device__ __forceinline__ void add(double var, double innov, double inv_var) {
        quad = fma(innov * innov, inv_var, quad);
        if (var > 1e-290 && var < 1e290) {
            int e;
            prod *= mantissa_and_exponent(var, &e);
            esum += e;
            if (++n_in_prod == 512) {  // prod < 2^512: renormalise well before overflow
                int e2;
                prod = mantissa_and_exponent(prod, &e2);
                esum += e2;
                n_in_prod = 0;
            }
        } else {
            logsum += loglik_slow_log(var);  // NaN for var < 0 or NaN, -inf for 0: same class as the reference
        }
    }
    __device__ __forceinline__ double value() const {
        return -0.5 * (log(prod) + (double)esum * 0.693147180559945309417232121458 + logsum) - 0.5 * quad;
    }
};

// Run the recursion over `len` staged points (dt, y, next-point yerr^2); the last `len - nadv`
// (0 or 1) points are only scored, not advanced past (end of the light curve).
// PF = true: the three operands of step i+1 are loaded while step i is computed (software prefetch);
// used when the series is read straight from global memory (K4, K5), pointless for shared memory.
template <int P, bool ALLC, bool PF = false>
__device__ __forceinline__ void filter_span_impl(KalmanReal<P>& kf, LogLikAcc& acc, const RealParams<P>& prm,
                                                 const double* __restrict__ sdt, const double* __restrict__ sy,
                                                 const double* __restrict__ se, int len, int nadv) {
    double y_c = 0.0, dt_c = 0.0, e_c = 0.0;
    if (PF && len > 0) { y_c = sy[0]; dt_c = sdt[0]; e_c = se[0]; }
    for (int i = 0; i < nadv; i++) {
        double y_i, dt_i, e_i;
        if (PF) {
            y_i = y_c; dt_i = dt_c; e_i = e_c;
            const int j = i + 1;  // j <= nadv <= len - 1 whenever the last point is only scored
            if (j < len) { y_c = sy[j]; if (j < nadv) { dt_c = sdt[j]; e_c = se[j]; } }
        } else {
            y_i = sy[i]; dt_i = sdt[i]; e_i = se[i];
        }
        double innov = (y_i - prm.mu) - kf.mean;
        double inv = rcp_fast(kf.var);
        acc.add(kf.var, innov, inv);
        kf.template advance<ALLC>(prm, innov, inv, dt_i, e_i);
    }
    if (nadv < len) {
        double y_l = PF ? y_c : sy[len - 1];
        double innov = (y_l - prm.mu) - kf.mean;
        double inv = rcp_fast(kf.var);
        acc.add(kf.var, innov, inv);
    }
}

template <int P, bool PF>
__device__ __forceinline__ void filter_span_any(KalmanReal<P>& kf, LogLikAcc& acc, const RealParams<P>& prm,
                                                const double* __restrict__ sdt, const double* __restrict__ sy,
                                                const double* __restrict__ se, int len, int nadv) {
    constexpr unsigned ALL = (P / 2 > 0) ? ((1u << (P / 2)) - 1u) : 0u;
    // Warp-uniform choice: the straight-line all-conjugate-pairs loop only when EVERY active lane of the
    // warp qualifies; otherwise all lanes run the generic loop (per-slot selection, reconverging each
    // slot).  A per-lane choice would execute both loops back to back in a mixed warp.
    const bool all_c = __all_sync(__activemask(), prm.cmask == ALL);
    if (all_c) filter_span_impl<P, true, PF>(kf, acc, prm, sdt, sy, se, len, nadv);
    else filter_span_impl<P, false, PF>(kf, acc, prm, sdt, sy, se, len, nadv);
}

template <int P>
__device__ __forceinline__ void filter_span(KalmanReal<P>& kf, LogLikAcc& acc, const RealParams<P>& prm,
                                            const double* __restrict__ sdt, const double* __restrict__ sy,
                                            const double* __restrict__ se, int len, int nadv) {
    filter_span_any<P, false>(kf, acc, prm, sdt, sy, se, len, nadv);
}

}  // namespace carma
