// mcmc.cu -- K3: persistent parallel-tempering MCMC kernel.  Everything of one MCMC run happens in
// ONE kernel launch: starting values, robust-adaptive-Metropolis proposals (Student-t_8 draws from
// Philox), the Kalman log-density of every proposal, accept/reject, the Robbins-Monro rank-1
// Cholesky adaptation, the tempered-chain exchanges and the storage of the coolest chain.  There is
// no host round trip per iteration.
//
// Reference behaviour restated (paths relative to /root/reference/src):
//   RunCarmaSampler / RunCar1Sampler   carmcmc.cpp:30-177   ladder, initial proposal covariance, step order
//   Sampler::Run / Iterate             samplers.cpp:37-115  start values, burn-in, thinning, sample storage
//   AdaptiveMetro::DoStep / Accept     steps.cpp:36-107
//   CholUpdateR1                       steps.cpp:111-131
//   ExchangeStep::DoStep               steps.hpp:318-362
//   CARp/CARMA/ZCARMA/CAR1::StartingValue, StartingAR, StartingMA   carpack.cpp:38-83, 175-230, 268-311,
//                                                                   416-477, 515-519, 586-644, 681-684
//
// Thread mapping: one thread = one chain = (ensemble, temperature).  All chains of an ensemble sit in
// one block; the light curve is staged once per block in shared memory (TMA bulk copy) and stays
// resident for the whole run.
//
// Step order.  The reference runs, per iteration, RAM(T-1), X(T-1,T-2), RAM(T-2), ..., X(1,0), RAM(0)
// sequentially (carmcmc.cpp:147-157).  order_mode 0 reproduces exactly that dependency graph as a
// skewed pipeline: at tick k chain i performs its RAM step of iteration n = k - (T-1-i); at the end of
// the tick the exchanges X(c,c-1) of the chains that just stepped are applied from the coldest pair up.
// Every RAM/exchange step sees precisely the state it would see in the sequential order, and all T
// filters of an ensemble run concurrently.  Random numbers are addressed, not consumed, so the
// evaluation order does not matter.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "kalman_real.cuh"
#include "series.h"

namespace carma {

constexpr int PT_BLOCK = 64;
constexpr size_t PT_SMEM_MAX = 96 * 1024;  // series resident in shared memory up to this size
constexpr size_t PT_SMEM_HELP_MAX = 200 * 1024;  // helper mode: one block per SM, ring + parameters on top

struct PTParams {
    int kind, q, d;
    carma_prior_t prior;
    int nsamples, burnin, thin, T, total_iters;
    double tmax;
    int dof;
    double target, gamma;
    unsigned long long seed;
    unsigned ens_offset;
    int max_start, order_mode, record;
    unsigned long long n_ens;
    // series statistics for the starting values
    double y_mean, y_var_sample, y_var_pop, median_dt, tspan;
    double dt_max;       // longest gap of the series (all curves in multi mode): rate clamp of transform_theta
    int ny;
    int series_in_smem;  // 1: sdt/sy/se of the log-density calls point into shared memory
    int pipelined;       // 1: software-pipelined filter loop (few warps per SM: latency bound)
    int r_in_smem;       // helper mode: the Cholesky factors of the block's chains live in shared memory
    // device buffers
    const double* init;  // d values or nullptr
    double* samples;     // [n_ens][nsamples][d]
    double* logposts;    // [n_ens][nsamples]
    double* accept_rates;    // [n_ens][T] or nullptr
    double* exchange_rates;  // [n_ens][T] or nullptr
    double* chol;            // [ntri][nthreads_total] packed upper factors
    carma_pt_trace_rec_t* ram_trace;   // [n_ens][iters][T]
    carma_pt_trace_rec_t* exch_trace;  // [n_ens][iters][T]
    double* proposals;                 // [n_ens][iters][T][d]
    int* status;                       // !=0 : a chain found no finite starting value
    // time-sliced mode (slice_ticks > 0): the launch has fewer blocks ("workers") than block-sized groups of
    // ensembles; a worker pulls (group, slice of ticks) units from a queue, slice-major, and the chains' state is
    // parked in HBM between slices.  Keeps every SM sub-partition at the same number of resident warps when the
    // number of groups is not a multiple of what the GPU holds, and any number of ensembles in ONE wave.
    int slice_ticks;
    unsigned n_groups;
    int* slice_queue;                  // next unit
    int* slice_done;                   // [n_groups] slices completed
    double* slice_state;               // [MAX_D + 1][nthreads_total]: theta, log-posterior
    int* slice_counts;                 // [3][nthreads_total]: accepted, exchanges tried, exchanges accepted
};

// multi light-curve mode: every block works on ONE curve of a ragged batch (its own series, prior and
// starting-value statistics); ensembles are numbered globally curve-major.
struct PTMulti {
    const double* dt;
    const double* y;
    const double* e2;
    const long long* off;
    const CurveInfo* info;
    int blocks_per_curve;
    int max_nyp;
    int enabled;
    unsigned long long ens_per_curve;
    int resident;  // 1: the series is staged in shared memory; 0: too long, read from global memory (L1/L2)
};

// ---- starting-value RNG (mirrors oracle StartRng draw for draw) ---------------------------------
struct StartRng {
    unsigned long long seed;
    uint32_t chain, attempt, blk;
    __device__ void next2(double* u0, double* u1) { uniforms2(seed, chain, STREAM_START, attempt, blk++, u0, u1); }
    __device__ double uniform() { double a, b; next2(&a, &b); return a; }
    __device__ double normal() { double a, b; next2(&a, &b); return normal_from(a, b); }
    __device__ double chisqr(int dof) {
        double acc = 0.0, prod = 1.0;
        int pairs = dof / 2, inprod = 0;
        for (int i = 0; i < pairs; i += 2) {
            double a, b;
            next2(&a, &b);
            prod *= a; inprod++;
            if (i + 1 < pairs) { prod *= b; inprod++; }
            if (inprod >= 8) { acc += -2.0 * log(prod); prod = 1.0; inprod = 0; }
        }
        if (inprod > 0) acc += -2.0 * log(prod);
        if (dof & 1) { double z = normal(); acc += z * z; }
        return acc;
    }
    __device__ double scaled_inverse_chisqr(int dof, double ssqr) { return ssqr / chisqr(dof) * (double)dof; }
};

// carpack.cpp:268-311
template <int P>
__device__ void starting_ar(StartRng& g, const PTParams& pp, double* loga) {
    constexpr int NL = (P + 1) / 2;
    constexpr double PI = 3.14159265358979323846;
    const double min_freq = 1.0 / pp.tspan;
    const double lr = log(pp.prior.max_freq / min_freq), l0 = log(min_freq);
    double cent[NL], width[NL];
    for (int i = 0; i < NL; i++) cent[i] = exp(lr * g.uniform() + l0);
    for (int i = 1; i < NL; i++) {  // sort descending
        double v = cent[i];
        int j = i - 1;
        while (j >= 0 && cent[j] < v) { cent[j + 1] = cent[j]; j--; }
        cent[j + 1] = v;
    }
    for (int i = 0; i < NL; i++) width[i] = exp(lr * g.uniform() + l0);
    if (P & 1) {
        cent[P / 2] = 0.0;
        double hi = (P / 2 >= 1) ? log(cent[(P / 2 >= 1) ? P / 2 - 1 : 0]) : log(pp.prior.max_freq);
        width[P / 2] = exp(l0 + (hi - l0) * g.uniform());
    }
    for (int i = 0; i < P / 2; i++) {
        double re = -2.0 * PI * width[i], im = 2.0 * PI * cent[i];
        loga[2 * i] = log(re * re + im * im);
        loga[2 * i + 1] = log(-2.0 * re);
    }
    if (P & 1) loga[P - 1] = log(2.0 * PI * width[P / 2]);
}

// __noinline__: one copy of the filter loop per kernel (three call sites), with its own register
// allocation, so the MCMC bookkeeping around it does not inflate the per-thread register count.
template <int P>
__device__ __noinline__ double logdensity_resident(const PTParams& pp, const MathTab& tb, const double* th, const double* sdt,
                                                   const double* sy, const double* se, double e2_0) {
    RealParams<P> prm;
    if (transform_theta<P>(pp.kind, pp.q, 0u, pp.prior, th, pp.dt_max, prm) != TT_OK) return -INFINITY;
    KalmanReal<P> kf;
    LogLikAcc acc;
    kf.reset(prm, e2_0);
    acc.init();
    const SeriesPtr gsrc{sdt, sy, se};
    if (pp.series_in_smem) {
        // the staged series: LDS off one 32-bit address register
        const uint32_t a = smem_u32(sdt);
        const SeriesSmem src{a, smem_u32(sy) - a, smem_u32(se) - a};
        if (pp.pipelined) filter_span_any_pipelined<P>(kf, acc, prm, tb, src, pp.ny, pp.ny - 1);
        else filter_span_any<P, false>(kf, acc, prm, tb, src, pp.ny, pp.ny - 1);
    } else {
        filter_span_any<P, true>(kf, acc, prm, tb, gsrc, pp.ny, pp.ny - 1);
    }
    if (acc.bad()) return loglik_exact_slow<P>(prm, tb, gsrc, pp.ny, e2_0) + prm.logprior;
    return acc.value() + prm.logprior;
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-specialised evaluation for SMALL launches (one ensemble = the reference's own use, or the 100 short runs that
// seed get_mle): with one or two warps per SM nothing hides latency, and a lone warp needs ~1,150 cycles per Kalman
// step although it only issues ~300 instructions.  More than half of those instructions -- the transition blocks
// exp(omega dt) of every slot -- depend on (theta, dt) alone, not on the filter state.  In helper mode a block carries,
// for each of its two chain warps, TWO producer warps (even / odd steps) that compute these factors ahead of the
// recursion and hand them over through a ring in shared memory; the chain warp runs only the state recursion
// (reciprocal, gain, covariance update, propagation, observation: ~135 instructions per step).  Hand-over is by
// monotone step counters in shared memory (volatile, one writer each), checked once per HELP_CHUNK steps.
// Same operations on the same values: results are bit-identical to the plain kernel.
// ---------------------------------------------------------------------------------------------------------------
constexpr int HELP_RING = 16;    // steps held in the ring, per chain thread
constexpr int HELP_CHUNK = 4;    // steps per hand-over check
constexpr int HELP_GROUPS = 2;   // producer warps per chain warp

template <int P>
struct HelpShared {
    static constexpr int NS = P / 2;
    static constexpr int NF = 2 * NS + 1;            // fa[NS], fb[NS], fo
    double* ring;                                    // [HELP_RING][NF][PT_BLOCK]
    double* par;                                     // [P][PT_BLOCK]: le[(P+1)/2], ls[P/2]
    int* pari;                                       // [PT_BLOCK]: cmask | active << 16
    volatile int* posted;                            // [2]  evaluations published by chain warp w
    volatile int* skip;                              // [2]  1: no lane of warp w evaluates in this tick
    volatile int* produced;                          // [HELP_GROUPS][2]  steps (absolute count) produced so far
    volatile int* consumed;                          // [2]
    volatile int* zready;                            // [2]  ticks whose random numbers are in zbuf
    double* zbuf;                                    // [2 (tick parity)][ZROWS][PT_BLOCK]: t draws, accept and exchange uniforms
    static constexpr int ZROWS = MAX_D + 2;
    __host__ __device__ static size_t doubles() {
        return (size_t)HELP_RING * NF * PT_BLOCK + (size_t)P * PT_BLOCK + PT_BLOCK / 2 + 8 + (size_t)2 * ZROWS * PT_BLOCK;
    }
    __device__ void carve(double* base) {
        ring = base;
        par = ring + (size_t)HELP_RING * NF * PT_BLOCK;
        pari = (int*)(par + (size_t)P * PT_BLOCK);
        int* flags = pari + PT_BLOCK;
        posted = flags; skip = flags + 2; produced = flags + 4; consumed = flags + 8; zready = flags + 10;
        zbuf = par + (size_t)P * PT_BLOCK + PT_BLOCK / 2 + 8;
    }
};

// Barrier of the chain threads around the exchanges of a tick.  In the warp-specialised kernel only the PT_BLOCK chain
// threads take part (named barrier with an explicit count): the producer warps are ordered by the hand-over flags
// alone and never reach a block-wide barrier inside the tick loop.
template <bool HELP>
__device__ __forceinline__ void tick_barrier() {
    if (HELP) asm volatile("bar.sync 1, %0;" ::"n"(PT_BLOCK) : "memory");
    else __syncthreads();
}

__device__ __forceinline__ void spin_until(volatile int* flag, int target) {
    while ((int)(*flag - target) < 0) {
    }
    __threadfence_block();
}

// chain side: every lane of the chain warp calls this together; `want` = this lane evaluates theta `th`
template <int P>
__device__ __noinline__ double logdensity_assisted(const PTParams& pp, const MathTab& tb, const HelpShared<P>& hs, bool want,
                                                   const double* th, const double* sy, const double* se, double e2_0,
                                                   int eval_idx) {
    constexpr int NS = P / 2, NF = HelpShared<P>::NF;
    const int t64 = threadIdx.x, w = t64 >> 5, lane = t64 & 31;
    RealParams<P> prm;
    bool act = want;
    if (act) act = transform_theta<P>(pp.kind, pp.q, 0u, pp.prior, th, pp.dt_max, prm) == TT_OK;
    if (act) {
#pragma unroll
        for (int k = 0; k < (P + 1) / 2; k++) hs.par[(size_t)k * PT_BLOCK + t64] = prm.le[k];
#pragma unroll
        for (int k = 0; k < NS; k++) hs.par[(size_t)((P + 1) / 2 + k) * PT_BLOCK + t64] = prm.ls[k];
    }
    hs.pari[t64] = act ? (int)(prm.cmask | (1u << 16)) : 0;
    const bool any = __any_sync(0xffffffffu, act);
    __threadfence_block();
    __syncwarp();
    const int nadv = pp.ny - 1;
    const int base = eval_idx * nadv;
    if (lane == 0) { hs.skip[w] = any ? 0 : 1; hs.consumed[w] = base; __threadfence_block(); hs.posted[w] = eval_idx + 1; }
    if (!any) return -INFINITY;
    KalmanReal<P> kf;
    LogLikAcc acc;
    if (act) kf.reset(prm, e2_0);
    acc.init();
    const uint32_t ya = smem_u32(sy), ea = smem_u32(se), ra = smem_u32(hs.ring) + 8u * (uint32_t)t64;
    int since = 0;
    for (int i0 = 0; i0 < nadv; i0 += HELP_CHUNK) {
        const int i1 = min(nadv, i0 + HELP_CHUNK);
        // the factors of steps i0 .. i1-1: even steps come from group 0, odd ones from group 1
        const int last = i1 - 1;
        const int last_even = (last & 1) ? last - 1 : last, last_odd = (last & 1) ? last : last - 1;   // i0 is even
        spin_until(&hs.produced[0 * 2 + w], base + last_even + 1);
        if (last_odd >= i0) spin_until(&hs.produced[1 * 2 + w], base + last_odd + 1);
        if (act) {
            for (int i = i0; i < i1; i++) {
                double fa[NS > 0 ? NS : 1], fb[NS > 0 ? NS : 1], fsb[NS > 0 ? NS : 1], fo, y_i, e_i;
                const uint32_t slot = ra + (uint32_t)(i % HELP_RING) * (uint32_t)(NF * PT_BLOCK * 8);
#pragma unroll
                for (int k = 0; k < NS; k++) {
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(fa[k]) : "r"(slot + (uint32_t)(k * PT_BLOCK * 8)));
                    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(fb[k]) : "r"(slot + (uint32_t)((NS + k) * PT_BLOCK * 8)));
                    fsb[k] = ((prm.cmask >> k) & 1u) ? -fb[k] : fb[k];
                }
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(fo) : "r"(slot + (uint32_t)(2 * NS * PT_BLOCK * 8)));
                asm("ld.shared.f64 %0, [%1];" : "=d"(y_i) : "r"(ya + 8u * (uint32_t)i));
                asm("ld.shared.f64 %0, [%1];" : "=d"(e_i) : "r"(ea + 8u * (uint32_t)i));
                const double innov = (y_i - prm.mu) - kf.mean;
                const double inv = rcp_fast(kf.var);
                acc.add(kf.var, innov, inv);
                kf.measurement_update(innov, inv);
                kf.propagate(prm, fa, fb, fsb, fo, e_i);
            }
            since += i1 - i0;
            if (since >= RENORM_EVERY - HELP_CHUNK) { acc.renorm(since); since = 0; }
        }
        __syncwarp();
        if (lane == 0) { __threadfence_block(); hs.consumed[w] = base + i1; }
    }
    if (!act) return -INFINITY;
    {
        double y_l;
        asm("ld.shared.f64 %0, [%1];" : "=d"(y_l) : "r"(ya + 8u * (uint32_t)nadv));
        const double innov = (y_l - prm.mu) - kf.mean;
        const double inv = rcp_fast(kf.var);
        acc.add(kf.var, innov, inv);
        acc.renorm(since + 1);
    }
    if (acc.bad()) return NAN;   // marker: the caller re-evaluates this lane with logdensity_resident (exact slow path inside)
    return acc.value() + prm.logprior;
}

// producer side: one warp of group g serving chain warp w, for every tick of the run
template <int P>
__device__ __noinline__ void helper_loop(const PTParams& pp, const MathTab& tb, const HelpShared<P>& hs, int g, int w,
                                         const double* sdt, int nticks) {
    constexpr int NS = P / 2, NF = HelpShared<P>::NF, ZR = HelpShared<P>::ZROWS;
    const int lane = threadIdx.x & 31, t64 = w * 32 + lane;
    const int nadv = pp.ny - 1;
    const uint32_t da = smem_u32(sdt), ra = smem_u32(hs.ring) + 8u * (uint32_t)t64;
    // the chain this lane serves (same numbering as pt_kernel)
    const int T = pp.T, epb = PT_BLOCK / T, e_local = t64 / T, ci = t64 % T;
    const unsigned long long ens = (unsigned long long)blockIdx.x * epb + e_local;
    const bool chain_active = (e_local < epb) && (ens < pp.n_ens);
    const uint32_t chain = (uint32_t)((pp.ens_offset + ens) * (unsigned long long)T + ci);
    // group 0 also draws the random numbers of the NEXT tick while the chain warp is still filtering: they are
    // addressed by (seed, chain, iteration, slot), not consumed from a state, so they can be made ahead of time
    auto draw_tick = [&](int tick) {
        const int n = pp.order_mode == 0 ? tick - (T - 1 - ci) : tick;
        if (chain_active && n >= 0 && n < pp.total_iters) {
            double* zb = hs.zbuf + (size_t)(tick & 1) * ZR * PT_BLOCK + t64;
            for (int j = 0; j < pp.d; j++) zb[(size_t)j * PT_BLOCK] = tdist_draw(pp.seed, chain, STREAM_PROPOSAL, (uint32_t)n, (uint32_t)j, pp.dof);
            double u0, u1;
            uniforms2(pp.seed, chain, STREAM_ACCEPT, (uint32_t)n, 0u, &u0, &u1);
            zb[(size_t)MAX_D * PT_BLOCK] = u0;
            uniforms2(pp.seed, chain, STREAM_EXCHANGE, (uint32_t)n, 0u, &u0, &u1);
            zb[(size_t)(MAX_D + 1) * PT_BLOCK] = u0;
        }
        __syncwarp();
        if (lane == 0) { __threadfence_block(); hs.zready[w] = tick + 1; }
    };
    if (g == 0) draw_tick(0);
    for (int e = 0; e < nticks; e++) {
        spin_until(&hs.posted[w], e + 1);
        if (!hs.skip[w]) {
            const int pi = hs.pari[t64];
            const bool act = (pi >> 16) & 1;
            RealParams<P> prm;
            prm.cmask = (unsigned)(pi & 0xffff);
            if (act) {
#pragma unroll
                for (int k = 0; k < (P + 1) / 2; k++) prm.le[k] = hs.par[(size_t)k * PT_BLOCK + t64];
#pragma unroll
                for (int k = 0; k < NS; k++) prm.ls[k] = hs.par[(size_t)((P + 1) / 2 + k) * PT_BLOCK + t64];
            }
            constexpr unsigned ALL = (NS > 0) ? ((1u << NS) - 1u) : 0u;
            const bool all_c = __all_sync(0xffffffffu, !act || prm.cmask == ALL);
            const int base = e * nadv;
            for (int i = g; i < nadv; i += HELP_GROUPS) {
                spin_until(&hs.consumed[w], base + i + 1 - HELP_RING);   // slot i % RING was used by step i - RING
                if (act) {
                    double dt, fa[NS > 0 ? NS : 1], fb[NS > 0 ? NS : 1], fsb[NS > 0 ? NS : 1], fo;
                    asm("ld.shared.f64 %0, [%1];" : "=d"(dt) : "r"(da + 8u * (uint32_t)i));
                    if (all_c) KalmanReal<P>::template transition<true>(prm, tb, dt, fa, fb, fsb, &fo);
                    else KalmanReal<P>::template transition<false>(prm, tb, dt, fa, fb, fsb, &fo);
                    const uint32_t slot = ra + (uint32_t)(i % HELP_RING) * (uint32_t)(NF * PT_BLOCK * 8);
#pragma unroll
                    for (int k = 0; k < NS; k++) {
                        asm volatile("st.shared.f64 [%0], %1;" ::"r"(slot + (uint32_t)(k * PT_BLOCK * 8)), "d"(fa[k]) : "memory");
                        asm volatile("st.shared.f64 [%0], %1;" ::"r"(slot + (uint32_t)((NS + k) * PT_BLOCK * 8)), "d"(fb[k]) : "memory");
                    }
                    asm volatile("st.shared.f64 [%0], %1;" ::"r"(slot + (uint32_t)(2 * NS * PT_BLOCK * 8)), "d"(fo) : "memory");
                }
                __syncwarp();
                if (lane == 0) { __threadfence_block(); hs.produced[g * 2 + w] = base + i + 1; }
            }
            // steps this group does not own must not leave the counter behind the other group's view of "all done"
            __syncwarp();
            if (lane == 0) { __threadfence_block(); hs.produced[g * 2 + w] = base + nadv; }
        }
        if (g == 0 && e + 1 < nticks) draw_tick(e + 1);
    }
}

template <int P>
__device__ double starting_value_attempt(const PTParams& pp, const MathTab& tb, StartRng& g, double* th, const double* sdt,
                                         const double* sy, const double* se, double e2_0) {
    const int n = pp.ny;
    if (pp.kind == CARMA_KIND_CAR1) {
        double sd = sqrt(g.scaled_inverse_chisqr(n - 1, pp.y_var_sample));
        double mu = pp.y_mean + (sd / (double)n) * g.normal();
        double log_omega = -log(pp.median_dt * (1.0 + 49.0 * g.uniform()));
        log_omega = fmin(log_omega, pp.prior.max_freq);  // sic (carpack.cpp:56)
        double scale = g.scaled_inverse_chisqr((int)pp.prior.measerr_dof, 1.0);
        scale = fmax(fmin(scale, 1.99), 0.51);
        th[0] = sd; th[1] = scale; th[2] = mu; th[3] = log_omega;
        return logdensity_resident<P>(pp, tb, th, sdt, sy, se, e2_0);
    }
    starting_ar<P>(g, pp, th + 3);
    if (pp.kind == CARMA_KIND_CARMA)
        for (int i = 0; i < pp.q; i++) th[3 + P + i] = fabs(g.normal());
    if (pp.kind == CARMA_KIND_ZCARMA) {
        double u = g.uniform();
        th[3 + P] = log(u / (1.0 - u));
    }
    double yvar = g.scaled_inverse_chisqr(n - 1, pp.y_var_sample);
    double mu = pp.y_mean + (sqrt(yvar) / (double)n) * g.normal();
    double scale = g.scaled_inverse_chisqr((int)pp.prior.measerr_dof, 1.0);
    scale = fmax(fmin(scale, 1.99), 0.51);
    th[0] = sqrt(yvar); th[1] = scale; th[2] = mu;
    return logdensity_resident<P>(pp, tb, th, sdt, sy, se, e2_0);
}

__device__ __forceinline__ int tri(int k, int j) { return j * (j + 1) / 2 + k; }  // k <= j

// min blocks/SM = 5 (<= 204 registers): the filter loop then keeps its whole state in registers (ncu r01d:
// with a 128-register cap the loop spilled 3 loads + 2 stores per step and stalled on them), and
// 5 x 148 = 740 resident blocks still hold BASELINE config 3 (683 blocks) in a single wave.
// HELP: warp-specialised variant for small launches (see HelpShared): threads [0, 64) are the chains, [64, 192) four
// producer warps; one block per SM.
template <int P, bool HELP>
__global__ void __launch_bounds__(HELP ? PT_BLOCK * (1 + HELP_GROUPS) : PT_BLOCK, HELP ? 1 : (P <= 5 ? 5 : 4))
pt_kernel(SeriesView sv, PTParams pp_in, size_t chol_stride, PTMulti mm) {
    extern __shared__ __align__(16) double smem[];
    __shared__ __align__(8) uint64_t bar;

    PTParams pp = pp_in;
    const int tid = HELP ? (threadIdx.x < PT_BLOCK ? threadIdx.x : 0) : threadIdx.x;   // helpers shadow chain thread 0 in the set-up code
    const bool is_helper = HELP && threadIdx.x >= PT_BLOCK;
    const int T = pp.T, d = pp.d;
    const int epb = PT_BLOCK / T;
    const int e_local = tid / T, i = tid % T;
    const bool sliced = !HELP && pp.slice_ticks > 0;
    size_t gtid = (size_t)blockIdx.x * PT_BLOCK + tid;
    unsigned long long ens;
    bool active;
    __shared__ int s_unit;
    const int nyp = mm.resident ? (mm.enabled ? mm.max_nyp : sv.nyp) : 0;

    // [ dt | y | e2n | exchange area ] (the math tables are static shared arrays)
    double* sdt = smem;
    double* sy = sdt + nyp;
    double* se = sdt + 2 * (size_t)nyp;
    double* xth = sdt + 3 * (size_t)nyp + 2;           // [PT_BLOCK][d] exchange area
    double* xlp = xth + (size_t)PT_BLOCK * d;           // [PT_BLOCK]
    double* xu = xlp + PT_BLOCK;                        // [PT_BLOCK] exchange uniforms
    double* xtemp = xu + PT_BLOCK;                      // [PT_BLOCK] temperature ladder (carmcmc.cpp:92-95), computed once
    double e2_0;
    MathTab tb;
    tb.load();  // ends with __syncthreads()
    if (mm.enabled) {
        // ---- this block's curve: coalesced cooperative copy (ragged offsets are not 16-byte aligned,
        // so no bulk copy here); se is the yerr^2 array shifted by one point
        const int curve = blockIdx.x / mm.blocks_per_curve;
        const unsigned long long e_in = (unsigned long long)(blockIdx.x % mm.blocks_per_curve) * epb + e_local;
        active = !is_helper && (e_local < epb) && (e_in < mm.ens_per_curve);
        ens = (unsigned long long)curve * mm.ens_per_curve + e_in;
        const long long o0 = mm.off[curve];
        const int ny = (int)(mm.off[curve + 1] - o0);
        const CurveInfo ci = mm.info[curve];
        pp.prior = ci.prior;
        pp.y_mean = ci.y_mean; pp.y_var_sample = ci.y_var_sample; pp.y_var_pop = ci.y_var_pop;
        pp.median_dt = ci.median_dt; pp.tspan = ci.tspan; pp.ny = ny;
        if (mm.resident) {
            for (int k = threadIdx.x; k < ny; k += blockDim.x) {
                sdt[k] = mm.dt[o0 + k];
                sy[k] = mm.y[o0 + k];
                se[k] = mm.e2[o0 + k];
            }
            __syncthreads();
        } else {
            sdt = const_cast<double*>(mm.dt + o0);
            sy = const_cast<double*>(mm.y + o0);
            se = const_cast<double*>(mm.e2 + o0);
        }
        e2_0 = se[0];
        se = se + 1;
    } else {
        ens = (unsigned long long)blockIdx.x * epb + e_local;
        active = !is_helper && (e_local < epb) && (ens < pp.n_ens);
        // ---- stage the light curve once (TMA bulk copy), resident for the whole run; a series that does
        // not fit in shared memory is read from global memory instead (every lane reads the same address:
        // one L1 line per warp, and the filter loop prefetches one step ahead)
        if (!mm.resident) {
            sdt = const_cast<double*>(sv.dt);
            sy = const_cast<double*>(sv.y);
            se = const_cast<double*>(sv.e2n);
        }
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0 && mm.resident) {
            // pieces of at most 64 KiB keep every transaction count far below the mbarrier tx limit
            const uint32_t total = (uint32_t)(3 * (size_t)sv.nyp * 8);
            mbar_expect_tx(&bar, total);
            const uint32_t piece = 1u << 16;
            for (uint32_t o = 0; o < total; o += piece) {
                uint32_t b = min(piece, total - o);
                bulk_g2s((char*)smem + o, (const char*)sv.dt + o, b, &bar);
            }
        }
        if (mm.resident) mbar_wait(&bar, 0);
        e2_0 = sv.e2_0;
    }

    const double temp = (T > 1) ? exp(log(pp.tmax) * (double)i / (double)(T - 1)) : 1.0;  // carmcmc.cpp:92-95
    if (!is_helper && tid < T) xtemp[tid] = temp;   // threads 0..T-1 are the chains of the block's first ensemble: i == tid
    const int total = pp.total_iters;
    const int nticks = pp.order_mode == 0 ? total + T - 1 : total;
    const int n_slices = sliced ? (nticks + pp.slice_ticks - 1) / pp.slice_ticks : 1;

  // ---- one pass per unit of work: the whole run of this block's ensembles, or (time-sliced mode) one slice of
  // ticks of the group of ensembles pulled from the queue
  for (;;) {
    int tick_begin = 0, tick_end = nticks;
    if (sliced) {
        if (threadIdx.x == 0) s_unit = atomicAdd(pp.slice_queue, 1);
        __syncthreads();
        const unsigned unit = (unsigned)s_unit;
        __syncthreads();
        if (unit >= pp.n_groups * (unsigned)n_slices) break;
        const unsigned grp = unit % pp.n_groups;
        const int slice = (int)(unit / pp.n_groups);
        tick_begin = slice * pp.slice_ticks;
        tick_end = min(nticks, tick_begin + pp.slice_ticks);
        gtid = (size_t)grp * PT_BLOCK + tid;
        ens = (unsigned long long)grp * epb + e_local;
        active = (e_local < epb) && (ens < pp.n_ens);
        if (slice > 0) {
            // the previous slice of this group was pulled n_groups units ago, normally long finished
            if (threadIdx.x == 0) spin_until((volatile int*)(pp.slice_done + grp), slice);
            __syncthreads();
            __threadfence();
        }
    }
    const uint32_t chain = (uint32_t)((pp.ens_offset + ens) * (unsigned long long)T + i);

    double th[MAX_D];
    double lp = -INFINITY;
#pragma unroll
    for (int j = 0; j < MAX_D; j++) th[j] = 0.0;
    int naccept = 0, nx_try = 0, nx_acc = 0;

    // ---- initial proposal Cholesky factor (carmcmc.cpp:127-136; diagonal, so R = sqrt(diag)).  Packed upper
    // triangle, [entry][chain]: in HBM (L2-resident) for full launches, in shared memory in helper mode when it fits
    // (a lone warp pays the full L2 latency on every one of the d(d+1)/2 dependent loads of the rank-1 update)
    double* R = pp.chol + gtid;
    if (HELP && pp.r_in_smem) {
        R = xtemp + PT_BLOCK + HelpShared<P>::doubles() + tid;
        chol_stride = PT_BLOCK;
    }
    // sliced mode: another SM may have written this group's factor since this SM last read it -> bypass L1
    auto rload = [&](const double* q) { return sliced ? __ldcg(q) : *q; };
    if (active && tick_begin > 0) {
        for (int j = 0; j < d; j++) th[j] = __ldcg(pp.slice_state + (size_t)j * chol_stride + gtid);
        lp = __ldcg(pp.slice_state + (size_t)MAX_D * chol_stride + gtid);
        naccept = __ldcg(pp.slice_counts + gtid);
        nx_try = __ldcg(pp.slice_counts + chol_stride + gtid);
        nx_acc = __ldcg(pp.slice_counts + 2 * chol_stride + gtid);
    }
    if (active && tick_begin == 0) {
        for (int j = 0; j < d; j++)
            for (int k = 0; k <= j; k++) R[(size_t)tri(k, j) * chol_stride] = 0.0;
        for (int j = 0; j < d; j++) R[(size_t)tri(j, j) * chol_stride] = 0.01;
        R[(size_t)tri(0, 0) * chol_stride] = sqrt(2.0 * pp.y_var_pop * pp.y_var_pop / (double)pp.ny);
        R[(size_t)tri(2, 2) * chol_stride] = sqrt(pp.y_var_pop / (double)pp.ny);
    }

    HelpShared<P> hs{};
    if (HELP) {
        hs.carve(xtemp + PT_BLOCK);
        if (threadIdx.x < 12) ((int*)hs.posted)[threadIdx.x] = 0;   // posted[2], skip[2], produced[4], consumed[2], zready[2]
        __syncthreads();
        if (is_helper) {
            const int hw = (threadIdx.x - PT_BLOCK) >> 5;            // helper warp 0..3: group = hw >> 1, chain warp = hw & 1
            helper_loop<P>(pp, tb, hs, hw >> 1, hw & 1, sdt, nticks);
            return;
        }
    }

    // ---- starting values (samplers.cpp:75-93)
    if (active && tick_begin == 0) {
        bool ok = false;
        if (pp.init) {
            for (int j = 0; j < d; j++) th[j] = pp.init[j];
            lp = logdensity_resident<P>(pp, tb, th, sdt, sy, se, e2_0);
            ok = isfinite(lp);
        }
        if (!ok) {
            for (int a = 0; a < pp.max_start && !ok; a++) {
                StartRng g{pp.seed, chain, (uint32_t)a, 0u};
                lp = starting_value_attempt<P>(pp, tb, g, th, sdt, sy, se, e2_0);
                ok = isfinite(lp);
            }
        }
        if (!ok) atomicExch(pp.status, 1);
    }

    for (int tick = tick_begin; tick < tick_end; tick++) {
        const int n = pp.order_mode == 0 ? tick - (T - 1 - i) : tick;
        const bool stepping = active && n >= 0 && n < total;
        double z[MAX_D], sp[MAX_D], nv[MAX_D];
        double znorm2 = 0.0, lpn = -INFINITY;
        double u_acc = 0.0, u_exch = 0.0;   // helper mode: the tick's uniforms, drawn ahead by a producer warp
        if (HELP) spin_until(&hs.zready[tid >> 5], tick + 1);
        if (stepping) {
            // ---- AdaptiveMetro::DoStep (steps.cpp:60-107)
            if (HELP) {
                const double* zb = hs.zbuf + (size_t)(tick & 1) * HelpShared<P>::ZROWS * PT_BLOCK + tid;
                for (int j = 0; j < d; j++) { z[j] = zb[(size_t)j * PT_BLOCK]; znorm2 += z[j] * z[j]; }
                u_acc = zb[(size_t)MAX_D * PT_BLOCK];
                u_exch = zb[(size_t)(MAX_D + 1) * PT_BLOCK];
            } else {
                for (int j = 0; j < d; j++) {
                    z[j] = tdist_draw(pp.seed, chain, STREAM_PROPOSAL, (uint32_t)n, (uint32_t)j, pp.dof);
                    znorm2 += z[j] * z[j];
                }
            }
            for (int j = 0; j < d; j++) {  // chol_factor_.t() * unit_proposal
                double s = 0.0;
                for (int k = 0; k <= j; k++) s += rload(R + (size_t)tri(k, j) * chol_stride) * z[k];
                sp[j] = s;
                nv[j] = th[j] + s;
            }
            for (int j = d; j < MAX_D; j++) nv[j] = 0.0;
        }
        if (HELP) {
            // every lane of the chain warps takes part in the hand-over protocol; `stepping` lanes evaluate
            lpn = logdensity_assisted<P>(pp, tb, hs, stepping, nv, sy, se, e2_0, tick);
            if (stepping && lpn != lpn) lpn = logdensity_resident<P>(pp, tb, nv, sdt, sy, se, e2_0);
        } else if (stepping) {
            lpn = logdensity_resident<P>(pp, tb, nv, sdt, sy, se, e2_0);
        }
        if (stepping) {
            // ---- Accept (steps.cpp:36-56)
            double alpha = (lpn - lp) / temp;
            double u = NAN;
            bool acc = false;
            if (!isfinite(alpha)) {
                alpha = 0.0;
            } else {
                if (HELP) {
                    u = u_acc;
                } else {
                    double u1;
                    uniforms2(pp.seed, chain, STREAM_ACCEPT, (uint32_t)n, 0u, &u, &u1);
                }
                alpha = fmin(exp(alpha), 1.0);
                if (u < alpha) { acc = true; naccept++; }
            }
            if (pp.record) {
                size_t r = ((size_t)ens * total + n) * T + i;
                carma_pt_trace_rec_t rec;
                rec.lp_prop = lpn; rec.lp_cur = lp; rec.alpha = alpha; rec.u = u; rec.accepted = acc; rec.pad = 0;
                pp.ram_trace[r] = rec;
                for (int j = 0; j < d; j++) pp.proposals[r * d + j] = nv[j];
            }
            if (acc) {
                for (int j = 0; j < d; j++) th[j] = nv[j];
                lp = lpn;
            }
            // ---- scale-matrix adaptation while niter < burnin (steps.cpp:82-99)
            if (n < pp.burnin) {
                double step = fmin(1.0, (double)d / pow((double)n, pp.gamma));
                double f = sqrt(step * fabs(alpha - pp.target)) / sqrt(znorm2);
                const double sign = (alpha < pp.target) ? -1.0 : 1.0;
                for (int j = 0; j < d; j++) sp[j] = f * sp[j];
                for (int k = 0; k < d; k++) {  // CholUpdateR1 (steps.cpp:111-131)
                    double lkk = rload(R + (size_t)tri(k, k) * chol_stride);
                    double r = sqrt(lkk * lkk + sign * sp[k] * sp[k]);
                    double c = r / lkk;
                    double s = sp[k] / lkk;
                    R[(size_t)tri(k, k) * chol_stride] = r;
                    for (int j = k + 1; j < d; j++) {
                        double lkj = (rload(R + (size_t)tri(k, j) * chol_stride) + sign * s * sp[j]) / c;
                        R[(size_t)tri(k, j) * chol_stride] = lkj;
                        sp[j] = c * sp[j] - s * lkj;
                    }
                }
            }
            // exchange uniform of ExchangeStep(i) at this iteration (drawn unconditionally, steps.hpp:337)
            if (i > 0) {
                double u0 = u_exch, u1;
                if (!HELP) uniforms2(pp.seed, chain, STREAM_EXCHANGE, (uint32_t)n, 0u, &u0, &u1);
                xu[tid] = u0;
            }
            // order_mode 0: chain 0 finished iteration n -> store before this tick's exchanges
            if (pp.order_mode == 0 && i == 0 && n >= pp.burnin && ((n - pp.burnin + 1) % pp.thin) == 0) {
                int sidx = (n - pp.burnin + 1) / pp.thin - 1;
                if (sidx < pp.nsamples) {
                    size_t o = (size_t)ens * pp.nsamples + sidx;
                    for (int j = 0; j < d; j++) pp.samples[o * d + j] = th[j];
                    pp.logposts[o] = lp;
                }
            }
        }
        if (T > 1) {
            // ---- exchanges (steps.hpp:318-362) through shared memory
            if (active) {
                for (int j = 0; j < d; j++) xth[(size_t)tid * d + j] = th[j];
                xlp[tid] = lp;
            }
            tick_barrier<HELP>();
            if (active && i == 0) {
                const int base = tid;  // thread of chain 0 of this ensemble
                for (int s = 1; s < T; s++) {
                    // order_mode 0: coldest pair first; order_mode 1: hottest first (reference order)
                    const int c = pp.order_mode == 0 ? s : T - s;
                    const int nc = pp.order_mode == 0 ? tick - (T - 1 - c) : tick;
                    if (nc < 0 || nc >= total) continue;
                    const double tc = xtemp[c], tcm = xtemp[c - 1];
                    const double this_lp = xlp[base + c], other_lp = xlp[base + c - 1];
                    double a = 1.0 / tc * (other_lp - this_lp) + 1.0 / tcm * (this_lp - other_lp);
                    // std::min(exp(a), 1.0) keeps a NaN (steps.hpp:334-337 then sets alpha = 0); CUDA's fmin would
                    // return the non-NaN operand, i.e. 1 -> always swap: test before the clamp
                    const double ea = exp(a);
                    a = (ea == ea) ? fmin(ea, 1.0) : 0.0;
                    const double ux = xu[base + c];
                    const bool sw = ux < a;
                    if (sw) {
                        for (int j = 0; j < d; j++) {
                            double tmp = xth[(size_t)(base + c) * d + j];
                            xth[(size_t)(base + c) * d + j] = xth[(size_t)(base + c - 1) * d + j];
                            xth[(size_t)(base + c - 1) * d + j] = tmp;
                        }
                        xlp[base + c] = other_lp;
                        xlp[base + c - 1] = this_lp;
                    }
                    if (pp.record) {
                        size_t r = ((size_t)ens * total + nc) * T + c;
                        carma_pt_trace_rec_t rec;
                        rec.lp_prop = other_lp; rec.lp_cur = this_lp; rec.alpha = a; rec.u = ux; rec.accepted = sw; rec.pad = 0;
                        pp.exch_trace[r] = rec;
                    }
                    // per-pair counters live with chain 0's thread: pack into the shared area afterwards
                    xu[base + c] = sw ? 2.0 : 3.0;  // consumed marker: 2 = swapped, 3 = tried
                }
            }
            tick_barrier<HELP>();
            if (active) {
                for (int j = 0; j < d; j++) th[j] = xth[(size_t)tid * d + j];
                lp = xlp[tid];
                if (i > 0 && stepping) {
                    double m = xu[tid];
                    if (m == 2.0) { nx_try++; nx_acc++; }
                    else if (m == 3.0) nx_try++;
                }
            }
        }
        if (pp.order_mode == 1 && stepping && i == 0 && n >= pp.burnin && ((n - pp.burnin + 1) % pp.thin) == 0) {
            int sidx = (n - pp.burnin + 1) / pp.thin - 1;
            if (sidx < pp.nsamples) {
                size_t o = (size_t)ens * pp.nsamples + sidx;
                for (int j = 0; j < d; j++) pp.samples[o * d + j] = th[j];
                pp.logposts[o] = lp;
            }
        }
    }

    if (active && tick_end == nticks) {
        if (pp.accept_rates) pp.accept_rates[(size_t)ens * T + i] = total > 0 ? (double)naccept / (double)total : 0.0;
        if (pp.exchange_rates) pp.exchange_rates[(size_t)ens * T + i] = nx_try > 0 ? (double)nx_acc / (double)nx_try : 0.0;
    }
    if (!sliced) break;
    // ---- park the chains and publish the slice
    if (active && tick_end < nticks) {
        for (int j = 0; j < d; j++) pp.slice_state[(size_t)j * chol_stride + gtid] = th[j];
        pp.slice_state[(size_t)MAX_D * chol_stride + gtid] = lp;
        pp.slice_counts[gtid] = naccept;
        pp.slice_counts[chol_stride + gtid] = nx_try;
        pp.slice_counts[2 * chol_stride + gtid] = nx_acc;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) atomicExch(pp.slice_done + (gtid / PT_BLOCK), tick_end / pp.slice_ticks + (tick_end == nticks ? 1 : 0));
  }
}

// One starting value per thread (chains chain0 .. chain0 + n - 1), the series read from global memory: the draw
// loop of pt_kernel on its own, for host-driven samplers (Parameter<>::StartingValue of the class API).
template <int P>
__global__ void start_value_kernel(SeriesView sv, PTParams pp, uint32_t chain0, int n, double* __restrict__ theta_out,
                                   double* __restrict__ lp_out, int* __restrict__ status) {
    MathTab tb;
    tb.load();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double th[MAX_D];
#pragma unroll
    for (int j = 0; j < MAX_D; j++) th[j] = 0.0;
    double lp = -INFINITY;
    bool ok = false;
    for (int a = 0; a < pp.max_start && !ok; a++) {
        StartRng g{pp.seed, chain0 + (uint32_t)k, (uint32_t)a, 0u};
        lp = starting_value_attempt<P>(pp, tb, g, th, sv.dt, sv.y, sv.e2n, sv.e2_0);
        ok = isfinite(lp);
    }
    if (!ok) atomicExch(status, 1);
    for (int j = 0; j < pp.d; j++) theta_out[(size_t)k * pp.d + j] = th[j];
    lp_out[k] = lp;
}

static size_t pt_smem_bytes(int nyp, int d) {
    return (3 * (size_t)nyp + 2 + (size_t)PT_BLOCK * d + 3 * PT_BLOCK) * sizeof(double);
}

// the dynamic shared memory opt-in of both kernel variants for every order at once, once per device:
// cudaFuncSetAttribute waits for running kernels, so a launch of a new order arriving while other runs are in flight
// (concurrent model fits draw their starts with short PT runs) would stall behind them (see OncePerDevice)
template <int P>
static cudaError_t pt_attr_from() {
    cudaError_t e = cudaFuncSetAttribute(pt_kernel<P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PT_SMEM_MAX);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(pt_kernel<P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PT_SMEM_HELP_MAX);
    if (e != cudaSuccess) return e;
    if constexpr (P < MAX_P) return pt_attr_from<P + 1>();
    else return cudaSuccess;
}
static cudaError_t pt_attrs() {
    static OncePerDevice once;
    return once.run([] { return pt_attr_from<1>(); });
}
template <int P>
static cudaError_t pt_plain_attr() { return pt_attrs(); }
template <int P>
static cudaError_t pt_help_attr() { return pt_attrs(); }

// Blocks of pt_kernel<P, false> one SM holds (registers + this launch's shared memory).
template <int P>
static int pt_blocks_per_sm(size_t smem) {
    int nb = 0;
    if (pt_plain_attr<P>() != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pt_kernel<P, false>, PT_BLOCK, smem) != cudaSuccess) return 0;
    return nb;
}

static int pt_blocks_per_sm(int p, size_t smem) {
    switch (p) {
        case 1: return pt_blocks_per_sm<1>(smem);
        case 2: return pt_blocks_per_sm<2>(smem);
        case 3: return pt_blocks_per_sm<3>(smem);
        case 4: return pt_blocks_per_sm<4>(smem);
        case 5: return pt_blocks_per_sm<5>(smem);
        case 6: return pt_blocks_per_sm<6>(smem);
        case 7: return pt_blocks_per_sm<7>(smem);
    }
    return 0;
}

// How many worker blocks a time-sliced launch should use for `groups` block-sized groups of ensembles, or 0 to
// launch one block per group.  Measured on B200 (scripts/pt_balance_probe.py, CARMA(5,3), ny = 1000): the time of
// a tick depends on the LARGEST number of warps any SM sub-partition hosts, t(1) : t(2) : t(3) = 1 : 1.37 : 1.79
// (two-warp blocks go to sub-partitions (0,1) and (2,3) alternately), so 683 groups on 148 SMs run at t(3)
// although two thirds of the sub-partitions hold two warps.  With W = 148 c workers (c even: every sub-partition
// holds c/2 warps) the run costs (groups / W) t(c/2).
static unsigned pt_slice_workers(unsigned groups, int sms, int occ) {
    if (occ < 1 || sms < 1) return 0;
    auto t = [](int k) { return 0.58 + 0.42 * k + (k > 3 ? 0.1 * (k - 3) : 0.0); };   // relative; k = warps per sub-partition
    double best;
    if (groups <= (unsigned)(sms * occ)) {
        const int b = (int)((groups + sms - 1) / sms);
        best = t((b + 1) / 2);
    } else {
        // more groups than the GPU holds: later waves run at whatever occupancy is left
        best = 0.0;
        for (unsigned left = groups; left > 0;) {
            const unsigned w = std::min(left, (unsigned)(sms * occ));
            best += t((int)(((w + sms - 1) / sms + 1) / 2));
            left -= w;
        }
    }
    unsigned pick = 0;
    for (int c = 2; c <= occ; c += 2) {
        const unsigned w = (unsigned)(sms * c);
        if (w >= groups) break;
        const double cost = (double)groups / (double)w * t(c / 2) * 1.02;   // 2 % for the queue and the parked state
        if (cost < best) { best = cost; pick = w; }
    }
    return pick;
}

template <int P>
static cudaError_t launch_pt(const SeriesView& sv, const PTParams& pp, size_t chol_stride, unsigned grid,
                             cudaStream_t stream, const PTMulti& mm, bool help) {
    size_t smem = pt_smem_bytes(mm.resident ? (mm.enabled ? mm.max_nyp : sv.nyp) : 0, pp.d);
    // always the same value (the residency rule keeps smem <= PT_SMEM_MAX): function attributes are process-wide,
    // and fits of different models launch this kernel concurrently from several host threads
    if (smem > PT_SMEM_MAX) return cudaErrorInvalidValue;
    if (help) {
        smem += HelpShared<P>::doubles() * sizeof(double);
        if (pp.r_in_smem) smem += (size_t)pp.d * (pp.d + 1) / 2 * PT_BLOCK * sizeof(double);
        cudaError_t e = pt_help_attr<P>();
        if (e != cudaSuccess) return e;
        pt_kernel<P, true><<<grid, PT_BLOCK * (1 + HELP_GROUPS), smem, stream>>>(sv, pp, chol_stride, mm);
        return cudaGetLastError();
    }
    cudaError_t e = pt_plain_attr<P>();
    if (e != cudaSuccess) return e;
    pt_kernel<P, false><<<grid, PT_BLOCK, smem, stream>>>(sv, pp, chol_stride, mm);
    return cudaGetLastError();
}

static bool valid_model_pt(int kind, int p, int q) {
    if (kind < CARMA_KIND_CAR1 || kind > CARMA_KIND_ZCARMA) return false;
    if (kind == CARMA_KIND_CAR1) return p == 1;
    if (p < 1 || p > MAX_P) return false;
    if (kind == CARMA_KIND_CARMA) return q >= 0 && q < p;
    return true;
}

}  // namespace carma

using namespace carma;

extern "C" {

void carma_pt_default_opts(carma_pt_opts_t* o) {
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->nsamples = 1000;
    o->burnin = 500;
    o->thin = 1;
    o->ntemps = 10;
    o->tmax = 100.0;
    o->dof = 8;
    o->target_rate = 0.25;
    o->gamma = 2.0 / 3.0;
    o->seed = 1;
    o->ensemble_offset = 0;
    o->max_start_attempts = 1000;
    o->order_mode = 0;
    o->record_trace = 0;
}

static int pt_check(carma_series_t s, int kind, int p, int q, const carma_prior_t* prior, const carma_pt_opts_t* o,
                    const void* samples, const void* logposts, const char* who) {
    if (!s || !prior || !o || !samples || !logposts) { set_error(std::string(who) + ": null argument"); return CARMA_ERR_ARG; }
    if (!valid_model_pt(kind, p, q)) { set_error(std::string(who) + ": invalid (kind,p,q)"); return CARMA_ERR_ARG; }
    if (o->ntemps < 1 || o->ntemps > PT_BLOCK || o->thin < 1 || o->nsamples < 0 || o->burnin < 0 || (o->dof & 1) || o->dof < 2) {
        set_error(std::string(who) + ": invalid options (1 <= ntemps <= 64, thin >= 1, even dof >= 2)");
        return CARMA_ERR_ARG;
    }
    return CARMA_OK;
}

// Fill PTParams, reserve the Cholesky scratch (owned by the series object) and launch.  All pointers
// are device pointers.  Exactly one of `s` / `m` is non-null.  *d_status_out receives the address of the
// device status word.
static int pt_launch(carma_series_t s, carma_multi_series_t m, const CurveInfo* d_info, int kind, int p, int q,
                     const carma_prior_t* prior, const carma_pt_opts_t* o, size_t n_ensembles, const double* d_init,
                     double* d_samples, double* d_logposts, double* d_accept_rates, double* d_exchange_rates,
                     carma_pt_trace_rec_t* d_rt, carma_pt_trace_rec_t* d_xt, double* d_pr, cudaStream_t st,
                     int** d_status_out) {
    SeriesView sv{};
    PTMulti mm{};
    PTParams pp{};
    pp.kind = kind; pp.q = q; pp.d = model_dim(kind, p, q);
    pp.nsamples = o->nsamples; pp.burnin = o->burnin; pp.thin = o->thin; pp.T = o->ntemps;
    pp.total_iters = o->burnin + o->nsamples * o->thin;
    pp.tmax = o->tmax; pp.dof = o->dof; pp.target = o->target_rate; pp.gamma = o->gamma;
    pp.seed = o->seed; pp.ens_offset = o->ensemble_offset; pp.max_start = std::max(1, o->max_start_attempts);
    pp.order_mode = o->order_mode; pp.record = (d_rt && d_xt && d_pr) ? 1 : 0;
    pp.init = d_init; pp.samples = d_samples; pp.logposts = d_logposts;
    pp.accept_rates = d_accept_rates; pp.exchange_rates = d_exchange_rates;
    pp.ram_trace = d_rt; pp.exch_trace = d_xt; pp.proposals = d_pr;
    const int epb = PT_BLOCK / pp.T;
    unsigned grid;
    DevBuf* scratch;
    if (s) {
        sv = s->view();
        pp.prior = *prior;
        pp.n_ens = n_ensembles;
        pp.y_mean = s->st.mean; pp.y_var_sample = s->st.var_sample; pp.y_var_pop = s->st.var_pop;
        pp.median_dt = s->st.median_dt; pp.tspan = s->st.tmax - s->st.tmin; pp.ny = (int)s->ny;
        pp.dt_max = s->dt_max;
        grid = (unsigned)((n_ensembles + epb - 1) / epb);
        scratch = &s->scratch_misc;
    } else {
        // n_ensembles = ensembles PER CURVE
        mm.enabled = 1;
        mm.dt = m->d_dt; mm.y = m->d_y; mm.e2 = m->d_e2; mm.off = m->d_off; mm.info = d_info;
        mm.blocks_per_curve = (int)((n_ensembles + epb - 1) / epb);
        mm.max_nyp = (m->max_ny + 1) & ~1;
        mm.ens_per_curve = n_ensembles;
        pp.n_ens = n_ensembles * m->ncurves;
        pp.dt_max = m->dt_max;
        grid = (unsigned)(m->ncurves * (size_t)mm.blocks_per_curve);
        scratch = &m->scratch_misc;
    }
    // series resident in shared memory when it fits (<= 96 KiB keeps at least two blocks per SM)
    mm.resident = pt_smem_bytes(mm.enabled ? mm.max_nyp : sv.nyp, pp.d) <= PT_SMEM_MAX ? 1 : 0;
    pp.series_in_smem = mm.resident;
    bool help = false;
    {
        // latency-bound regime (at most ~2 blocks per SM): pipelined loop; CARMA_PT_PIPE=0/1 overrides (measurements)
        const char* e = getenv("CARMA_PT_PIPE");
        pp.pipelined = e ? (e[0] == '1') : (grid <= 2u * 148u);
        // at most one block per SM, series resident, one series: warp-specialised kernel (CARMA_PT_HELP=0/1 overrides)
        const char* h = getenv("CARMA_PT_HELP");
        const size_t help_bytes = ((size_t)HELP_RING * (2 * (p / 2) + 1) * PT_BLOCK + (size_t)p * PT_BLOCK + PT_BLOCK / 2 + 8 +
                                   (size_t)2 * (MAX_D + 2) * PT_BLOCK) * sizeof(double);
        const bool can = !mm.enabled && mm.resident && pp.ny >= 2 && pt_smem_bytes(sv.nyp, pp.d) + help_bytes <= PT_SMEM_HELP_MAX;
        help = can && (h ? (h[0] == '1') : (grid <= 148u));
        const size_t r_bytes = (size_t)pp.d * (pp.d + 1) / 2 * PT_BLOCK * sizeof(double);
        pp.r_in_smem = (help && pt_smem_bytes(sv.nyp, pp.d) + help_bytes + r_bytes <= PT_SMEM_HELP_MAX) ? 1 : 0;
    }
    size_t nthreads = (size_t)grid * PT_BLOCK;
    size_t ntri = (size_t)pp.d * (pp.d + 1) / 2;
    // time-sliced launch (one series, not in helper mode): CARMA_PT_SLICE=0 disables, =W forces W workers
    unsigned workers = 0;
    const int nticks = pp.order_mode == 0 ? pp.total_iters + pp.T - 1 : pp.total_iters;
    if (!mm.enabled && !help && nticks >= 8) {
        const char* e = getenv("CARMA_PT_SLICE");
        if (e) {
            workers = (unsigned)std::max(0, atoi(e));
        } else {
            int dev = 0, sms = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            workers = pt_slice_workers(grid, sms, pt_blocks_per_sm(p, pt_smem_bytes(mm.resident ? sv.nyp : 0, pp.d)));
        }
        if (workers >= grid) workers = 0;
    }
    // [ chol | status, queue | done[groups] | parked state | parked counters ]
    const size_t off_flags = ntri * nthreads * sizeof(double);
    const size_t off_done = off_flags + 16;
    const size_t off_state = (off_done + (workers ? (size_t)grid * sizeof(int) : 0) + 15) & ~(size_t)15;
    const size_t off_counts = off_state + (workers ? (size_t)(MAX_D + 1) * nthreads * sizeof(double) : 0);
    const size_t bytes = off_counts + (workers ? 3 * nthreads * sizeof(int) : 0);
    if (!scratch->reserve(bytes)) return CARMA_ERR_CUDA;
    pp.chol = (double*)scratch->p;
    pp.status = (int*)((char*)scratch->p + off_flags);
    if (d_status_out) *d_status_out = pp.status;
    if (!cuda_ok(cudaMemsetAsync(pp.status, 0, off_state - off_flags, st), "memset status")) return CARMA_ERR_CUDA;
    unsigned launch_grid = grid;
    if (workers) {
        pp.slice_ticks = std::max(4, nticks / 32);
        pp.n_groups = grid;
        pp.slice_queue = pp.status + 1;
        pp.slice_done = (int*)((char*)scratch->p + off_done);
        pp.slice_state = (double*)((char*)scratch->p + off_state);
        pp.slice_counts = (int*)((char*)scratch->p + off_counts);
        launch_grid = workers;
    }
    cudaError_t e;
    switch (p) {
        case 1: e = launch_pt<1>(sv, pp, nthreads, launch_grid, st, mm, help); break;
        case 2: e = launch_pt<2>(sv, pp, nthreads, launch_grid, st, mm, help); break;
        case 3: e = launch_pt<3>(sv, pp, nthreads, launch_grid, st, mm, help); break;
        case 4: e = launch_pt<4>(sv, pp, nthreads, launch_grid, st, mm, help); break;
        case 5: e = launch_pt<5>(sv, pp, nthreads, launch_grid, st, mm, help); break;
        case 6: e = launch_pt<6>(sv, pp, nthreads, launch_grid, st, mm, help); break;
        case 7: e = launch_pt<7>(sv, pp, nthreads, launch_grid, st, mm, help); break;
        default: e = cudaErrorInvalidValue;
    }
    if (!cuda_ok(e, "pt_kernel launch")) return CARMA_ERR_CUDA;
    return CARMA_OK;
}

int carma_starting_value(carma_series_t s, int kind, int p, int q, const carma_prior_t* prior, uint64_t seed,
                         uint32_t chain, int max_attempts, double* theta_out, double* logpost_out) {
    if (!s || !prior || !theta_out || !logpost_out) { set_error("carma_starting_value: null argument"); return CARMA_ERR_ARG; }
    if (!valid_model_pt(kind, p, q)) { set_error("carma_starting_value: invalid (kind,p,q)"); return CARMA_ERR_ARG; }
    if (!cuda_ok(cudaSetDevice(s->device), "cudaSetDevice")) return CARMA_ERR_CUDA;
    PTParams pp{};
    pp.kind = kind; pp.q = q; pp.d = model_dim(kind, p, q);
    pp.prior = *prior;
    pp.seed = seed; pp.max_start = std::max(1, max_attempts);
    pp.y_mean = s->st.mean; pp.y_var_sample = s->st.var_sample; pp.y_var_pop = s->st.var_pop;
    pp.median_dt = s->st.median_dt; pp.tspan = s->st.tmax - s->st.tmin; pp.ny = (int)s->ny;
    pp.dt_max = s->dt_max;
    pp.series_in_smem = 0;
    const size_t d = (size_t)pp.d;
    if (!s->scratch_out.reserve((d + 1) * sizeof(double) + 16)) return CARMA_ERR_CUDA;
    double* d_th = (double*)s->scratch_out.p;
    double* d_lp = d_th + d;
    int* d_status = (int*)(d_lp + 1);
    if (!cuda_ok(cudaMemset(d_status, 0, sizeof(int)), "memset status")) return CARMA_ERR_CUDA;
    SeriesView sv = s->view();
#define LAUNCH_SV(PP) start_value_kernel<PP><<<1, 32>>>(sv, pp, chain, 1, d_th, d_lp, d_status)
    switch (p) {
        case 1: LAUNCH_SV(1); break;
        case 2: LAUNCH_SV(2); break;
        case 3: LAUNCH_SV(3); break;
        case 4: LAUNCH_SV(4); break;
        case 5: LAUNCH_SV(5); break;
        case 6: LAUNCH_SV(6); break;
        default: LAUNCH_SV(7); break;
    }
#undef LAUNCH_SV
    if (!cuda_ok(cudaGetLastError(), "start_value_kernel launch")) return CARMA_ERR_CUDA;
    int status = 0;
    bool ok = cuda_ok(cudaMemcpy(theta_out, d_th, d * sizeof(double), cudaMemcpyDeviceToHost), "D2H theta") &&
              cuda_ok(cudaMemcpy(logpost_out, d_lp, sizeof(double), cudaMemcpyDeviceToHost), "D2H logpost") &&
              cuda_ok(cudaMemcpy(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost), "D2H status");
    if (!ok) return CARMA_ERR_CUDA;
    if (status) { set_error("carma_starting_value: no finite starting value within max_attempts"); return CARMA_ERR_START; }
    return CARMA_OK;
}

int carma_pt_run_dev(carma_series_t s, int kind, int p, int q, const carma_prior_t* prior,
                     const carma_pt_opts_t* o, size_t n_ensembles, const double* d_init, double* d_samples,
                     double* d_logposts, double* d_accept_rates, double* d_exchange_rates, void* stream) {
    int rc = pt_check(s, kind, p, q, prior, o, d_samples, d_logposts, "carma_pt_run_dev");
    if (rc) return rc;
    if (n_ensembles == 0) return CARMA_OK;
    if (!cuda_ok(cudaSetDevice(s->device), "cudaSetDevice")) return CARMA_ERR_CUDA;
    return pt_launch(s, nullptr, nullptr, kind, p, q, prior, o, n_ensembles, d_init, d_samples, d_logposts,
                     d_accept_rates, d_exchange_rates, nullptr, nullptr, nullptr, (cudaStream_t)stream, nullptr);
}

int carma_pt_run(carma_series_t s, int kind, int p, int q, const carma_prior_t* prior, const carma_pt_opts_t* o,
                 size_t n_ensembles, const double* init, double* samples, double* logposts, double* accept_rates,
                 double* exchange_rates, carma_pt_trace_rec_t* ram_trace, carma_pt_trace_rec_t* exch_trace,
                 double* proposals) {
    int rc = pt_check(s, kind, p, q, prior, o, samples, logposts, "carma_pt_run");
    if (rc) return rc;
    if (n_ensembles == 0) return CARMA_OK;
    if (!cuda_ok(cudaSetDevice(s->device), "cudaSetDevice")) return CARMA_ERR_CUDA;
    // the run lives on the series' own non-blocking stream: runs driven from different host threads on different
    // series handles (choose_order) overlap instead of queueing on the default stream
    if (!s->slot_stream[1] && !cuda_ok(cudaStreamCreateWithFlags(&s->slot_stream[1], cudaStreamNonBlocking), "cudaStreamCreate")) return CARMA_ERR_CUDA;
    cudaStream_t st = s->slot_stream[1];
    const bool rec = o->record_trace && ram_trace && exch_trace && proposals;
    const size_t d = (size_t)model_dim(kind, p, q), T = (size_t)o->ntemps;
    const size_t iters = (size_t)o->burnin + (size_t)o->nsamples * o->thin;
    const size_t n_s = n_ensembles * (size_t)o->nsamples;
    const size_t n_tr = rec ? n_ensembles * iters * T : 0;
    size_t bytes_out = (n_s * d + n_s + 2 * n_ensembles * T) * sizeof(double);
    size_t bytes_tr = 2 * n_tr * sizeof(carma_pt_trace_rec_t) + n_tr * d * sizeof(double);
    if (!s->scratch_out.reserve(bytes_out + bytes_tr + 64) || !s->scratch_in.reserve(d * sizeof(double))) return CARMA_ERR_CUDA;
    double* d_samples = (double*)s->scratch_out.p;
    double* d_lp = d_samples + n_s * d;
    double* d_ar = d_lp + n_s;
    double* d_xr = d_ar + n_ensembles * T;
    carma_pt_trace_rec_t* d_rt = (carma_pt_trace_rec_t*)(d_xr + n_ensembles * T);
    carma_pt_trace_rec_t* d_xt = d_rt + n_tr;
    double* d_pr = (double*)(d_xt + n_tr);
    if (!cuda_ok(cudaMemsetAsync(s->scratch_out.p, 0, bytes_out + bytes_tr, st), "memset outputs")) return CARMA_ERR_CUDA;
    const double* d_init = nullptr;
    if (init) {
        if (!cuda_ok(cudaMemcpyAsync(s->scratch_in.p, init, d * sizeof(double), cudaMemcpyHostToDevice, st), "H2D init")) return CARMA_ERR_CUDA;
        d_init = (const double*)s->scratch_in.p;
    }
    int* d_status = nullptr;
    rc = pt_launch(s, nullptr, nullptr, kind, p, q, prior, o, n_ensembles, d_init, d_samples, d_lp, d_ar, d_xr,
                   rec ? d_rt : nullptr, rec ? d_xt : nullptr, rec ? d_pr : nullptr, st, &d_status);
    if (rc) return rc;
    int status = 0;
    if (!cuda_ok(cudaMemcpyAsync(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost, st), "D2H status")) return CARMA_ERR_CUDA;
    if (!cuda_ok(cudaStreamSynchronize(st), "pt_kernel")) return CARMA_ERR_CUDA;
    if (status) { set_error("carma_pt_run: a chain found no finite starting value within max_start_attempts"); return CARMA_ERR_START; }
    bool ok = cuda_ok(cudaMemcpyAsync(samples, d_samples, n_s * d * sizeof(double), cudaMemcpyDeviceToHost, st), "D2H samples") &&
              cuda_ok(cudaMemcpyAsync(logposts, d_lp, n_s * sizeof(double), cudaMemcpyDeviceToHost, st), "D2H logposts");
    if (ok && accept_rates) ok = cuda_ok(cudaMemcpyAsync(accept_rates, d_ar, n_ensembles * T * sizeof(double), cudaMemcpyDeviceToHost, st), "D2H accept");
    if (ok && exchange_rates) ok = cuda_ok(cudaMemcpyAsync(exchange_rates, d_xr, n_ensembles * T * sizeof(double), cudaMemcpyDeviceToHost, st), "D2H exch");
    if (ok && rec) {
        ok = cuda_ok(cudaMemcpyAsync(ram_trace, d_rt, n_tr * sizeof(carma_pt_trace_rec_t), cudaMemcpyDeviceToHost, st), "D2H ram trace") &&
             cuda_ok(cudaMemcpyAsync(exch_trace, d_xt, n_tr * sizeof(carma_pt_trace_rec_t), cudaMemcpyDeviceToHost, st), "D2H exch trace") &&
             cuda_ok(cudaMemcpyAsync(proposals, d_pr, n_tr * d * sizeof(double), cudaMemcpyDeviceToHost, st), "D2H proposals");
    }
    if (ok) ok = cuda_ok(cudaStreamSynchronize(st), "pt_run D2H");
    return ok ? CARMA_OK : CARMA_ERR_CUDA;
}

// One PT-MCMC run per light curve of a ragged batch (n_ensembles independent ensembles each), one launch.
int carma_multi_pt_run(carma_multi_series_t m, int kind, int p, int q, const carma_prior_t* priors,
                       const carma_pt_opts_t* o, size_t n_ensembles, double* samples, double* logposts,
                       double* accept_rates, double* exchange_rates) {
    if (!m || !o || !samples || !logposts) { set_error("carma_multi_pt_run: null argument"); return CARMA_ERR_ARG; }
    if (!valid_model_pt(kind, p, q)) { set_error("carma_multi_pt_run: invalid (kind,p,q)"); return CARMA_ERR_ARG; }
    if (o->ntemps < 1 || o->ntemps > PT_BLOCK || o->thin < 1 || o->nsamples < 0 || o->burnin < 0 || (o->dof & 1) || o->dof < 2) {
        set_error("carma_multi_pt_run: invalid options (1 <= ntemps <= 64, thin >= 1, even dof >= 2)");
        return CARMA_ERR_ARG;
    }
    if (n_ensembles == 0) return CARMA_OK;
    if (!cuda_ok(cudaSetDevice(m->device), "cudaSetDevice")) return CARMA_ERR_CUDA;
    const size_t d = (size_t)model_dim(kind, p, q), T = (size_t)o->ntemps, nc = m->ncurves;
    const size_t n_tot = nc * n_ensembles, n_s = n_tot * (size_t)o->nsamples;
    std::vector<CurveInfo> info(m->info);
    if (priors)
        for (size_t c = 0; c < nc; c++) info[c].prior = priors[c];
    size_t bytes_out = (n_s * d + n_s + 2 * n_tot * T) * sizeof(double);
    if (!m->scratch_out.reserve(bytes_out + 64) || !m->scratch_pr.reserve(nc * sizeof(CurveInfo))) return CARMA_ERR_CUDA;
    if (!cuda_ok(cudaMemcpy(m->scratch_pr.p, info.data(), nc * sizeof(CurveInfo), cudaMemcpyHostToDevice), "H2D curve info")) return CARMA_ERR_CUDA;
    if (!cuda_ok(cudaMemset(m->scratch_out.p, 0, bytes_out), "memset outputs")) return CARMA_ERR_CUDA;
    double* d_samples = (double*)m->scratch_out.p;
    double* d_lp = d_samples + n_s * d;
    double* d_ar = d_lp + n_s;
    double* d_xr = d_ar + n_tot * T;
    int* d_status = nullptr;
    int rc = pt_launch(nullptr, m, (const CurveInfo*)m->scratch_pr.p, kind, p, q, nullptr, o, n_ensembles, nullptr,
                       d_samples, d_lp, d_ar, d_xr, nullptr, nullptr, nullptr, 0, &d_status);
    if (rc) return rc;
    if (!cuda_ok(cudaDeviceSynchronize(), "pt_kernel (multi)")) return CARMA_ERR_CUDA;
    int status = 0;
    if (!cuda_ok(cudaMemcpy(&status, d_status, sizeof(int), cudaMemcpyDeviceToHost), "D2H status")) return CARMA_ERR_CUDA;
    if (status) { set_error("carma_multi_pt_run: a chain found no finite starting value within max_start_attempts"); return CARMA_ERR_START; }
    bool ok = cuda_ok(cudaMemcpy(samples, d_samples, n_s * d * sizeof(double), cudaMemcpyDeviceToHost), "D2H samples") &&
              cuda_ok(cudaMemcpy(logposts, d_lp, n_s * sizeof(double), cudaMemcpyDeviceToHost), "D2H logposts");
    if (ok && accept_rates) ok = cuda_ok(cudaMemcpy(accept_rates, d_ar, n_tot * T * sizeof(double), cudaMemcpyDeviceToHost), "D2H accept");
    if (ok && exchange_rates) ok = cuda_ok(cudaMemcpy(exchange_rates, d_xr, n_tot * T * sizeof(double), cudaMemcpyDeviceToHost), "D2H exch");
    return ok ? CARMA_OK : CARMA_ERR_CUDA;
}

}  // extern "C"
