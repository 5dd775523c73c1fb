// fast_math.cuh -- branch-free FP64 exp / sincos / reciprocal for the Kalman time loop.
//
// Why not the CUDA math library here: inside the loop each step needs p complex exponentials
// rho = exp(omega dt) (kfilter.cpp:200) and one reciprocal.  The library versions are accurate but
// each carries range checks and slow-path calls (Payne-Hanek, denormal scaling, division fix-up):
// every branch splits the basic block, so ptxas cannot interleave the 3-5 independent polynomial
// chains of one step, and with ~3.5 warps per scheduler the dependent-FMA latency is exposed (ncu
// round 1a: stall_wait 1.69 per issue, FP64 pipe 49 % active).  These versions are straight-line code
// with ~1 ulp accuracy on the ranges the filter can produce, and coefficients come from the
// constant bank so they are FMA operands instead of per-iteration register moves.
//
//   exp_fast(x)      any finite x; flushes to 0 below exp(-708) (the library returns denormals there)
//   sincos_fast(x)   |x| < 2^51: two-term Cody-Waite reduction with FMA (exact product), fdlibm kernels
//   rcp_fast(x)      MUFU.RCP64H seed + two Newton steps; x normal (0 -> inf/NaN, like 1/x -> non-finite)
#pragma once
#include <cuda_runtime.h>

namespace carma {

// (e^r - 1 - r)/r^2 on |r| <= ln2/2, degree 9, Chebyshev-node interpolation computed with mpmath at 60
// digits (max relative error of the resulting e^r approximation 1.6e-17 before rounding).
static __constant__ double kExpQ[10] = {0.5000000000000001,     0.16666666666666669,    0.04166666666662413,
                                 0.008333333333330062,   0.0013888888917213717,  0.00019841269863053618,
                                 2.4801521295954376e-05, 2.7557268459997064e-06, 2.7620088445409746e-07,
                                 2.510038549551032e-08};
// fdlibm __kernel_sin / __kernel_cos coefficients (|r| <= pi/4)
static __constant__ double kSinC[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                                2.75573137070700676789e-06,  -2.50507602534068634195e-08, 1.58969099521155010221e-10};
static __constant__ double kCosC[6] = {4.16666666666666019037e-02,  -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                                -2.75573143513906633035e-07, 2.08757232129817482790e-09,  -1.13596475577881948265e-11};

__device__ __forceinline__ double exp_fast(double x) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: rint() through the adder
    double t = fma(x, 1.4426950408889634074, MAGIC);
    int n = __double2loint(t);
    double fn = t - MAGIC;
    double r = fma(fn, -6.93147180369123816490e-01, x);  // ln2 hi / lo (fdlibm split)
    r = fma(fn, -1.90821492927058770002e-10, r);
    double q = kExpQ[9];
    q = fma(q, r, kExpQ[8]);
    q = fma(q, r, kExpQ[7]);
    q = fma(q, r, kExpQ[6]);
    q = fma(q, r, kExpQ[5]);
    q = fma(q, r, kExpQ[4]);
    q = fma(q, r, kExpQ[3]);
    q = fma(q, r, kExpQ[2]);
    q = fma(q, r, kExpQ[1]);
    q = fma(q, r, kExpQ[0]);
    double p = fma(q, r, 1.0);
    p = fma(p, r, 1.0);
    // 2^n by exponent construction; results below the normal range are flushed to zero
    int nn = max(n, -1022);
    double s = __hiloint2double((nn + 1023) << 20, 0);
    double res = p * s;
    return (n < -1022) ? 0.0 : res;
}

__device__ __forceinline__ void sincos_fast(double x, double* sn, double* cs) {
    const double MAGIC = 6755399441055744.0;
    double t = fma(x, 0.63661977236758134308, MAGIC);
    int q = __double2loint(t);
    double fn = t - MAGIC;
    double r = fma(fn, -1.5707963267948965580, x);      // pi/2 = hi + lo, product exact inside the FMA
    r = fma(fn, -6.1232339957367658860e-17, r);
    double z = r * r;
    double sp = kSinC[5];
    sp = fma(sp, z, kSinC[4]);
    sp = fma(sp, z, kSinC[3]);
    sp = fma(sp, z, kSinC[2]);
    sp = fma(sp, z, kSinC[1]);
    sp = fma(sp, z, kSinC[0]);
    double cp = kCosC[5];
    cp = fma(cp, z, kCosC[4]);
    cp = fma(cp, z, kCosC[3]);
    cp = fma(cp, z, kCosC[2]);
    cp = fma(cp, z, kCosC[1]);
    cp = fma(cp, z, kCosC[0]);
    double s = fma(r * z, sp, r);                     // r + r^3 S(z)
    double c = fma(z * z, cp, fma(-0.5, z, 1.0));     // 1 - z/2 + z^2 C(z)
    double so = (q & 1) ? c : s;
    double co = (q & 1) ? s : c;
    *sn = (q & 2) ? -so : so;
    *cs = ((q + 1) & 2) ? -co : co;
}

__device__ __forceinline__ double rcp_fast(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

}  // namespace carma
