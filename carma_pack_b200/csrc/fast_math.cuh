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
//   exp_fast(x)      any finite x; 32-entry 2^(j/32) table (L1-resident) + degree-6 polynomial; flushes to 0
//                    below exp(-708) (the library returns denormals there)
//   sincos_fast(x)   |x| < 2^51: two-term Cody-Waite reduction with FMA (exact product), fdlibm kernels
//   rcp_fast(x)      MUFU.RCP64H seed + two Newton steps; x normal (0 -> inf/NaN, like 1/x -> non-finite)
#pragma once
#include <cuda_runtime.h>

namespace carma {

// fdlibm __kernel_sin / __kernel_cos coefficients (|r| <= pi/4)
static __constant__ double kSinC[6] = {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,
                                2.75573137070700676789e-06,  -2.50507602534068634195e-08, 1.58969099521155010221e-10};
static __constant__ double kCosC[6] = {4.16666666666666019037e-02,  -1.38888888888741095749e-03, 2.48015872894767294178e-05,
                                -2.75573143513906633035e-07, 2.08757232129817482790e-09,  -1.13596475577881948265e-11};

// 2^(j/32), j = 0..31 (mpmath, correctly rounded)
static __device__ const double kExp2Tab[32] = {
    1.0, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237, 1.0905077326652577, 1.1143867425958924,
    1.1387886347566916, 1.1637248587775775, 1.189207115002721, 1.215247359980469, 1.241857812073484,
    1.2690509571917332, 1.2968395546510096, 1.3252366431597413, 1.3542555469368927, 1.383909881963832,
    1.4142135623730951, 1.4451808069770467, 1.4768261459394993, 1.5091644275934228, 1.5422108254079407,
    1.5759808451078865, 1.6104903319492543, 1.645755478153965, 1.681792830507429, 1.718619298122478,
    1.7562521603732995, 1.7947090750031072, 1.8340080864093424, 1.8741676341103, 1.9152065613971474,
    1.9571441241754002};

// exp(x) = 2^n * 2^(j/32) * e^r with |r| <= ln2/64: degree-6 Taylor for e^r - 1 (error 3.5e-18), one
// L1-resident table read; 12 FP64 instructions instead of 16 for the polynomial-only version.
__device__ __forceinline__ double exp_fast(double x) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: rint() through the adder
    double t = fma(x, 46.166241308446828384, MAGIC);  // 32 / ln 2
    int k = __double2loint(t);
    double fk = t - MAGIC;
    double r = fma(fk, -6.93147180369123816490e-01 / 32.0, x);  // ln2/32 hi / lo (fdlibm split, exact scaling)
    r = fma(fk, -1.90821492927058770002e-10 / 32.0, r);
    double q = 1.0 / 720.0;
    q = fma(q, r, 1.0 / 120.0);
    q = fma(q, r, 1.0 / 24.0);
    q = fma(q, r, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    double pm1 = q * r;                       // e^r - 1
    double tj = __ldg(&kExp2Tab[k & 31]);
    double p = fma(tj, pm1, tj);              // 2^(j/32) e^r
    int n = k >> 5;
    // 2^n by exponent construction; results below the normal range are flushed to zero.  The decision is
    // taken on x itself, so arguments far below the range of the magic-number rounding (|x| > 2^51 ln2/32)
    // still give exactly 0 instead of garbage.
    int nn = min(max(n, -1022), 1023);
    double s = __hiloint2double((nn + 1023) << 20, 0);
    double res = p * s;
    return (x < -708.0) ? 0.0 : res;
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
    // quadrant signs through the integer pipe (an FP64 negation would cost a DADD each)
    *sn = __hiloint2double(__double2hiint(so) ^ ((q & 2) << 30), __double2loint(so));
    *cs = __hiloint2double(__double2hiint(co) ^ (((q + 1) & 2) << 30), __double2loint(co));
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
