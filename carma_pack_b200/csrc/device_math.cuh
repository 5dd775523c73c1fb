// device_math.cuh -- small FP64 / complex-FP64 helpers, Philox4x32-10, and the sm_100a
// bulk-copy (TMA 1-D) + mbarrier wrappers used to stage light curves in shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace carma {

// ------------------------------------------------------------------ complex double
struct cxd {
    double re, im;
};
__host__ __device__ __forceinline__ cxd cx(double r, double i = 0.0) { return cxd{r, i}; }
__host__ __device__ __forceinline__ cxd operator+(cxd a, cxd b) { return cxd{a.re + b.re, a.im + b.im}; }
__host__ __device__ __forceinline__ cxd operator-(cxd a, cxd b) { return cxd{a.re - b.re, a.im - b.im}; }
__host__ __device__ __forceinline__ cxd operator-(cxd a) { return cxd{-a.re, -a.im}; }
__host__ __device__ __forceinline__ cxd operator*(cxd a, cxd b) {
    return cxd{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
__host__ __device__ __forceinline__ cxd operator*(double s, cxd a) { return cxd{s * a.re, s * a.im}; }
__host__ __device__ __forceinline__ cxd conj(cxd a) { return cxd{a.re, -a.im}; }
__host__ __device__ __forceinline__ cxd cdiv(cxd a, cxd b) {
    // Smith's algorithm: robust against overflow of |b|^2
    if (fabs(b.re) >= fabs(b.im)) {
        double r = b.im / b.re, den = b.re + b.im * r;
        return cxd{(a.re + a.im * r) / den, (a.im - a.re * r) / den};
    } else {
        double r = b.re / b.im, den = b.re * r + b.im;
        return cxd{(a.re * r + a.im) / den, (a.im * r - a.re) / den};
    }
}
__host__ __device__ __forceinline__ double cabs_(cxd a) { return hypot(a.re, a.im); }

// ------------------------------------------------------------------ Philox4x32-10
// Counter-based RNG: the MCMC kernels address draws by (stream, chain, iteration, slot) so no
// generator state is carried and no host round trip is needed (replaces the global
// boost::random::mt19937 of src/random.cpp:20).
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint64_t seed, uint32_t out[4]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = 0xD2511F53ull * (uint64_t)c0;
        uint64_t p1 = 0xCD9E8D57ull * (uint64_t)c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 53-bit uniform in the open interval (0,1)
__host__ __device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}

enum { STREAM_PROPOSAL = 0, STREAM_ACCEPT = 1, STREAM_EXCHANGE = 2, STREAM_START = 3 };

__host__ __device__ __forceinline__ void uniforms2(uint64_t seed, uint32_t chain, uint32_t stream, uint32_t iter,
                                                   uint32_t blk, double* u0, double* u1) {
    uint32_t w[4];
    philox4x32_10(blk, iter, chain, stream, seed, w);
    *u0 = u53(w[0], w[1]);
    *u1 = u53(w[2], w[3]);
}

__host__ __device__ __forceinline__ double normal_from(double u0, double u1) {
    return sqrt(-2.0 * log(u0)) * cos(6.283185307179586476925286766559 * u1);
}

// Student-t, even dof: N(0,1)/sqrt(chi2/dof), chi2 = -2 ln(prod of dof/2 uniforms).
// Same distribution as StudentProposal(8,1) (src/carmcmc.cpp:139, src/random.cpp:158-164).
__host__ __device__ __forceinline__ double tdist_draw(uint64_t seed, uint32_t chain, uint32_t stream, uint32_t iter,
                                                      uint32_t j, int dof) {
    int nblk = 1 + (dof / 2 + 1) / 2;
    uint32_t base = j * (uint32_t)nblk;
    double u0, u1;
    uniforms2(seed, chain, stream, iter, base, &u0, &u1);
    double z = normal_from(u0, u1);
    double prod = 1.0;
    int need = dof / 2;
    for (int b = 1; b < nblk; b++) {
        uniforms2(seed, chain, stream, iter, base + b, &u0, &u1);
        if (need > 0) { prod *= u0; need--; }
        if (need > 0) { prod *= u1; need--; }
    }
    double chi2 = -2.0 * log(prod);
    return z / sqrt(chi2 / (double)dof);
}

#ifdef __CUDACC__
// ------------------------------------------------------------------ mbarrier + bulk copy (TMA 1-D)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk async copy (TMA engine, SASS UBLKCP); dst/src 16-byte aligned, bytes % 16 == 0
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Exponent/mantissa split used to turn sum(log(var_i)) into one log of a running product.
__device__ __forceinline__ double mantissa_and_exponent(double v, int* e) {
    int hi = __double2hiint(v), lo = __double2loint(v);
    *e = ((hi >> 20) & 0x7ff) - 1023;
    hi = (hi & 0x800fffff) | 0x3ff00000;
    return __hiloint2double(hi, lo);
}
#endif

}  // namespace carma
