// kalman_group.cuh -- EXPERIMENT (not part of the product library; see DESIGN.md "Single-ensemble latency").
// The CARMA Kalman recursion of ONE parameter set spread over a group of 16 lanes, tried as a latency form of
// the PT-MCMC kernel for runs with a single ensemble.  Same recursion as KalmanReal (kalman_real.cuh;
// reference: KalmanFilterp::Reset/Update, kfilter.cpp:138-215) up to the order of the sums; verified against
// the one-thread filter by lane_group_probe.cu.  Measured on B200: 610 cycles per Kalman step for a lone warp
// (CARMA(5,3)) against 700 (all conjugate pairs) / 860 (generic roots) for the one-thread form -- and no gain
// at all inside the PT kernel, where the ten chains of an ensemble share one SM and the lane-split form spends
// a full FP64 issue slot on 16 (of which 9 useful) lanes.
#pragma once
#include "../../carma_pack_b200/csrc/kalman_real.cuh"

namespace carma {

// ---- one chain on PT_GROUP = 16 lanes ---------------------------------------------------------------
// A warp holds ONE chain: lanes 16-31 run an exact clone of lanes 0-15 (same data, same instructions), so every
// shuffle uses the compile-time full-warp mask with width 16 and the two halves can never diverge.  (In the
// latency regime this kernel is for, a half-empty warp costs nothing: the FP64 pipe takes the same two cycles
// for 16 or 32 active lanes.)
// The real-basis state is cut into NU = ceil(P/2) units (a 2x2 slot per conjugate/real pair of roots, the
// odd real root padded to a slot whose second component is identically zero; NU <= 4 for P <= 7).  Lane
// (a, b) = (l >> 2, l & 3) of the group owns the 2x2 block D_ab of D = P - V and the transition blocks Phi_a,
// Phi_b.  Per Kalman step every lane does the work of ONE block instead of all NU^2:
//   D_ab <- Phi_a (D_ab - g_a g_b^T / var) Phi_b^T          local, 22 FP64 instructions
//   g_a   = h_a + sum_b D_ab c_b                            2 butterfly rounds over b   (lanes xor 1, 2)
//   g_b                                                     from lane (b, 0)
//   var   = sum_a c_a.g_a + e2,  mean = sum_a c_a.z_a       2 butterfly rounds over a   (lanes xor 4, 8)
// which cuts the instruction stream of a step from ~330 to ~150 warp instructions (the quantity that bounds
// a lone warp).  Same recursion as KalmanReal (kalman_real.cuh) up to the order of the sums; the transition
// blocks of step i+1 are evaluated while the dependent chain of step i is in flight.
constexpr int PT_GROUP = 16;
constexpr unsigned GROUP_MASK = 0xffffffffu;

struct UnitPar {
    double lam0, lam1, c0, c1, h0, h1;
    bool conj, pair, valid;  // conjugate pair / any pair (second component exists) / unit exists
};

template <int P>
__device__ __forceinline__ UnitPar unit_of(const RealParams<P>& prm, int u) {
    constexpr int NU = (P + 1) / 2;
    UnitPar r;
    r.valid = u < NU;
    r.pair = (2 * u + 1) < P;
    r.conj = r.pair && ((prm.cmask >> u) & 1u);
    const int i0 = r.valid ? 2 * u : 0, i1 = r.pair ? 2 * u + 1 : 0;
    r.lam0 = r.valid ? prm.lam[i0] : 0.0;
    r.c0 = r.valid ? prm.c[i0] : 0.0;
    r.h0 = r.valid ? prm.h[i0] : 0.0;
    r.lam1 = r.pair ? prm.lam[i1] : 0.0;
    r.c1 = r.pair ? prm.c[i1] : 0.0;
    r.h1 = r.pair ? prm.h[i1] : 0.0;
    return r;
}

struct Phi2 { double f00, f01, f10, f11; };

__device__ __forceinline__ Phi2 unit_phi(const UnitPar& u, double dt) {
    const double e0 = exp_fast(u.lam0 * dt), e1 = exp_fast(u.lam1 * dt);
    double sn, cs;
    sincos_fast(u.lam1 * dt, &sn, &cs);
    const double ec = e0 * cs, es = e0 * sn;
    Phi2 f;
    f.f00 = u.conj ? ec : (u.valid ? e0 : 0.0);
    f.f01 = u.conj ? -es : 0.0;
    f.f10 = u.conj ? es : 0.0;
    f.f11 = u.conj ? ec : (u.pair ? e1 : 0.0);
    return f;
}

// log-likelihood of the staged series for one parameter set, evaluated by the 16 lanes of a group
// (l16 = lane & 15); every lane returns the same value.
__device__ __forceinline__ double group_filter(const UnitPar ua, const UnitPar ub, double v0, double scale, double mu,
                                               const double* __restrict__ sdt, const double* __restrict__ sy,
                                               const double* __restrict__ se, double e2_0, int ny, int l16) {
    constexpr unsigned gmask = GROUP_MASK;
    const int a = l16 >> 2, b = l16 & 3;
    double D00 = 0.0, D01 = 0.0, D10 = 0.0, D11 = 0.0;
    double za0 = 0.0, za1 = 0.0;
    double ga0 = ua.h0, ga1 = ua.h1, gb0 = ub.h0, gb1 = ub.h1;
    double var = v0 + scale * e2_0, mean = 0.0;
    LogLikAcc acc;
    acc.init();
    const int nadv = ny - 1;
    Phi2 fb = unit_phi(ub, nadv > 0 ? sdt[0] : 0.0);
    double y_n = sy[0], e_n = nadv > 0 ? se[0] : 0.0;
#pragma unroll 2
    for (int i = 0; i < nadv; i++) {
        const double y_i = y_n, e_i = e_n;
        y_n = sy[i + 1];                      // i + 1 <= ny - 1
        e_n = se[min(i + 1, nadv - 1)];
        // transition blocks of this step: own Phi_b, Phi_a from the diagonal lane (a, a)
        const Phi2 cb = fb;
        Phi2 ca;
        ca.f00 = __shfl_sync(gmask, cb.f00, a * 5, PT_GROUP);
        ca.f01 = __shfl_sync(gmask, cb.f01, a * 5, PT_GROUP);
        ca.f10 = __shfl_sync(gmask, cb.f10, a * 5, PT_GROUP);
        ca.f11 = __shfl_sync(gmask, cb.f11, a * 5, PT_GROUP);
        // ... and the next step's, independent of the filter state (fills the latency of the chain below)
        fb = unit_phi(ub, sdt[min(i + 1, nadv - 1)]);

        const double innov = (y_i - mu) - mean;
        const double inv = rcp_fast(var);
        acc.add(var, innov, inv);
        // measurement update (kfilter.cpp:191-197); (g_a g_b) inv keeps D_ab = D_ba^T exactly
        const double w = innov * inv;
        za0 = fma(ga0, w, za0);
        za1 = fma(ga1, w, za1);
        D00 = fma(-(ga0 * gb0), inv, D00);
        D01 = fma(-(ga0 * gb1), inv, D01);
        D10 = fma(-(ga1 * gb0), inv, D10);
        D11 = fma(-(ga1 * gb1), inv, D11);
        // transition (kfilter.cpp:200-206)
        {
            const double u0 = za0, u1 = za1;
            za0 = fma(ca.f00, u0, ca.f01 * u1);
            za1 = fma(ca.f10, u0, ca.f11 * u1);
            const double m00 = fma(ca.f00, D00, ca.f01 * D10), m01 = fma(ca.f00, D01, ca.f01 * D11);
            const double m10 = fma(ca.f10, D00, ca.f11 * D10), m11 = fma(ca.f10, D01, ca.f11 * D11);
            D00 = fma(m00, cb.f00, m01 * cb.f01);
            D01 = fma(m00, cb.f10, m01 * cb.f11);
            D10 = fma(m10, cb.f00, m11 * cb.f01);
            D11 = fma(m10, cb.f10, m11 * cb.f11);
        }
        // g = D c + h: sum over the column units b
        double u0 = fma(D00, ub.c0, D01 * ub.c1), u1 = fma(D10, ub.c0, D11 * ub.c1);
        u0 += __shfl_xor_sync(gmask, u0, 1, PT_GROUP);
        u1 += __shfl_xor_sync(gmask, u1, 1, PT_GROUP);
        u0 += __shfl_xor_sync(gmask, u0, 2, PT_GROUP);
        u1 += __shfl_xor_sync(gmask, u1, 2, PT_GROUP);
        ga0 = ua.h0 + u0;
        ga1 = ua.h1 + u1;
        gb0 = __shfl_sync(gmask, ga0, b * 4, PT_GROUP);
        gb1 = __shfl_sync(gmask, ga1, b * 4, PT_GROUP);
        // predicted observation (kfilter.cpp:208-210): sum over the row units a
        double pv = fma(ua.c0, ga0, ua.c1 * ga1), pm = fma(ua.c0, za0, ua.c1 * za1);
        pv += __shfl_xor_sync(gmask, pv, 4, PT_GROUP);
        pm += __shfl_xor_sync(gmask, pm, 4, PT_GROUP);
        pv += __shfl_xor_sync(gmask, pv, 8, PT_GROUP);
        pm += __shfl_xor_sync(gmask, pm, 8, PT_GROUP);
        var = __dmul_rn(scale, e_i) + pv;
        mean = pm;
    }
    {
        const double innov = (y_n - mu) - mean;
        const double inv = rcp_fast(var);
        acc.add(var, innov, inv);
    }
    return acc.value();
}

}  // namespace carma
