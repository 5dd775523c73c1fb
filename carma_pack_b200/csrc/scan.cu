// scan.cu -- K5: temporally parallel (associative-scan) CARMA Kalman log-likelihood for ONE very long
// light curve (BASELINE config 5: ny = 10^6).  The reference filter is a strictly sequential
// recurrence over time (kfilter.hpp:126-132); every other configuration has abundant batch
// parallelism and uses the sequential-in-registers kernels, this one has none.
//
// Formulation: Sarkka & Garcia-Fernandez (2021) filtering elements a_k = (A, b, C, eta, J) in the
// real half of the rotated state space (theta_transform.cuh), with the associative operator
//     A_ij = A_j M A_i                 b_ij = A_j M (b_i + C_i eta_j) + b_j
//     C_ij = A_j M C_i A_j^T + C_j     eta_ij = A_i^T M^T (eta_j - J_j b_i) + eta_i
//     J_ij = A_i^T M^T J_j A_i + J_i    M = (I + C_i J_j)^{-1}
// The prefix a_0 (x) ... (x) a_k carries the filtered mean and covariance (b, C) after point k.
//
// Three passes, O(ny) work, O(ny / chunk) parallelism:
//   1. scan_reduce_kernel : one thread per chunk of `chunk` consecutive points folds its per-point
//      elements into one aggregate.  The right operand is always a single-point element (J_j = w w^T/S
//      is rank one), so M comes from Sherman-Morrison: no matrix inverse in this pass.
//   2. scan_block / scan_totals / scan_apply kernels : a two-level Kogge-Stone scan of the aggregates over many
//      blocks (general operator, Gauss-Jordan inverse with partial pivoting on a PxP matrix) that emits the
//      filtered state in front of every chunk.
//   3. scan_filter_kernel : one thread per chunk re-runs the ordinary sequential filter (KalmanReal,
//      the same code as K1) from that state and sums its chunk's log-likelihood terms.
//   4. scan_sum_kernel    : fixed-order reduction of the chunk sums (+ log-prior): deterministic.
#include <algorithm>
#include <cmath>
#include <string>

#include "kalman_real.cuh"
#include "series.h"

namespace carma {

template <int P>
struct ScanParams {
    RealParams<P> prm;
    double V[P][P];  // stationary covariance, real basis
    int status;      // TT_OK / TT_NEG_INF
};

template <int P>
struct ScanElem {
    double A[P][P], b[P], C[P][P], eta[P], J[P][P];
};
template <int P>
struct FiltState {
    double b[P], C[P][P];
};

template <int P>
__global__ void scan_params_kernel(int kind, int q, int d, unsigned flags, carma_prior_t prior, double dt_max,
                                   const double* __restrict__ theta, ScanParams<P>* __restrict__ out, int nrows) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    theta += (size_t)row * d;
    out += row;
    double th[MAX_D];
    for (int j = 0; j < MAX_D; j++) th[j] = (j < d) ? theta[j] : 0.0;
    double Vr[P * (P + 1) / 2];
    RealParams<P> prm;
    int st = transform_theta<P, true>(kind, q, flags, prior, th, dt_max, prm, Vr);
    out->status = st;
    out->prm = prm;
    int o = 0;
    for (int m = 0; m < P; m++)
        for (int n = m; n < P; n++) { out->V[m][n] = Vr[o]; out->V[n][m] = Vr[o]; o++; }
}

// the same for a model given by (sigsqr, omega, ma): the KalmanFilterp class API
template <int P>
__global__ void scan_params_explicit_kernel(ExplicitModel ex, double dt_max, ScanParams<P>* __restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double Vr[P * (P + 1) / 2];
    RealParams<P> prm;
    int st = explicit_constants<P, true>(ex, dt_max, prm, Vr);
    out->status = st;
    out->prm = prm;
    int o = 0;
    for (int m = 0; m < P; m++)
        for (int n = m; n < P; n++) { out->V[m][n] = Vr[o]; out->V[n][m] = Vr[o]; o++; }
}

// dense transition matrix Phi(dt) in the real basis: 2x2 blocks [[A, sB],[B, A]] (kalman_real.cuh)
template <int P>
__device__ void build_phi(const RealParams<P>& prm, const MathTab& tb, double dt, double F[P][P]) {
    for (int i = 0; i < P; i++)
        for (int j = 0; j < P; j++) F[i][j] = 0.0;
    constexpr int NS = P / 2;
    double fa[NS > 0 ? NS : 1], fb[NS > 0 ? NS : 1], fsb[NS > 0 ? NS : 1], fo;
    KalmanReal<P>::template transition<false>(prm, tb, dt, fa, fb, fsb, &fo);
    for (int s = 0; s < NS; s++) {
        F[2 * s][2 * s] = fa[s]; F[2 * s][2 * s + 1] = fsb[s];
        F[2 * s + 1][2 * s] = fb[s]; F[2 * s + 1][2 * s + 1] = fa[s];
    }
    if (P & 1) F[P - 1][P - 1] = fo;
}

// per-point quantities of point k >= 1 reached from k-1 by dt:  Q = V - F V F^T, w = F^T c, Qc, S, K
template <int P>
struct StepOps {
    double F[P][P], Q[P][P], w[P], Qc[P], K[P], S;
    __device__ void build(const ScanParams<P>& sp, const MathTab& tb, double dt, double r) {
        build_phi<P>(sp.prm, tb, dt, F);
        double FV[P][P];
        for (int i = 0; i < P; i++)
            for (int j = 0; j < P; j++) {
                double s = 0.0;
                for (int k = 0; k < P; k++) s = fma(F[i][k], sp.V[k][j], s);
                FV[i][j] = s;
            }
        for (int i = 0; i < P; i++)
            for (int j = 0; j < P; j++) {
                double s = 0.0;
                for (int k = 0; k < P; k++) s = fma(FV[i][k], F[j][k], s);
                Q[i][j] = sp.V[i][j] - s;
            }
        S = r;
        for (int i = 0; i < P; i++) {
            double a = 0.0, q = 0.0;
            for (int k = 0; k < P; k++) { a = fma(F[k][i], obs_c<P>(k), a); q = fma(Q[i][k], obs_c<P>(k), q); }
            w[i] = a;
            Qc[i] = q;
        }
        for (int i = 0; i < P; i++) S = fma(obs_c<P>(i), Qc[i], S);
        for (int i = 0; i < P; i++) K[i] = Qc[i] / S;
    }
};

// element of a single point k >= 1
template <int P>
__device__ void single_elem(const ScanParams<P>& sp, const StepOps<P>& o, double y, ScanElem<P>& e) {
    for (int i = 0; i < P; i++) {
        for (int j = 0; j < P; j++) {
            double kcF = 0.0;  // (K c^T F)[i][j] = K_i * (c^T F)_j = K_i * w_j
            kcF = o.K[i] * o.w[j];
            e.A[i][j] = o.F[i][j] - kcF;
            e.C[i][j] = o.Q[i][j] - o.K[i] * o.Qc[j];
            e.J[i][j] = o.w[i] * o.w[j] / o.S;
        }
        e.b[i] = o.K[i] * y;
        e.eta[i] = o.w[i] * y / o.S;
    }
}

// element of point 0: Kalman update of the stationary prior (A = 0, eta = 0, J = 0)
template <int P>
__device__ void first_elem(const ScanParams<P>& sp, double y, double r, ScanElem<P>& e) {
    double Vc[P], S = r;
    for (int i = 0; i < P; i++) {
        double s = 0.0;
        for (int k = 0; k < P; k++) s = fma(sp.V[i][k], obs_c<P>(k), s);
        Vc[i] = s;
    }
    for (int i = 0; i < P; i++) S = fma(obs_c<P>(i), Vc[i], S);
    for (int i = 0; i < P; i++) {
        for (int j = 0; j < P; j++) {
            e.A[i][j] = 0.0;
            e.J[i][j] = 0.0;
            e.C[i][j] = sp.V[i][j] - Vc[i] * Vc[j] / S;
        }
        e.b[i] = Vc[i] * y / S;
        e.eta[i] = 0.0;
    }
}

// acc <- acc (x) a_k with a_k a single-point element: Sherman-Morrison, no inverse
template <int P>
__device__ void fold_step(const StepOps<P>& o, double y, ScanElem<P>& a) {
    double cw[P], aw[P];
    double sp_ = o.S, wb = 0.0;
    for (int i = 0; i < P; i++) {
        double s1 = 0.0, s2 = 0.0;
        for (int k = 0; k < P; k++) { s1 = fma(a.C[i][k], o.w[k], s1); s2 = fma(a.A[k][i], o.w[k], s2); }
        cw[i] = s1;
        aw[i] = s2;
        wb = fma(o.w[i], a.b[i], wb);
    }
    for (int i = 0; i < P; i++) sp_ = fma(o.w[i], cw[i], sp_);
    const double inv = 1.0 / sp_;
    const double rr = (y - wb) * inv;
    double bt[P], Ct[P][P], At[P][P];
    for (int i = 0; i < P; i++) {
        a.eta[i] = fma(aw[i], rr, a.eta[i]);
        bt[i] = fma(cw[i], rr, a.b[i]);
        for (int j = 0; j < P; j++) {
            a.J[i][j] = fma(aw[i] * inv, aw[j], a.J[i][j]);
            Ct[i][j] = fma(-cw[i] * inv, cw[j], a.C[i][j]);
            At[i][j] = fma(-cw[i] * inv, aw[j], a.A[i][j]);
        }
    }
    // G = (I - K c^T) F
    double G[P][P];
    for (int i = 0; i < P; i++)
        for (int j = 0; j < P; j++) G[i][j] = o.F[i][j] - o.K[i] * o.w[j];
    double GC[P][P];
    for (int i = 0; i < P; i++) {
        double s = 0.0;
        for (int k = 0; k < P; k++) s = fma(G[i][k], bt[k], s);
        a.b[i] = fma(o.K[i], y, s);
        for (int j = 0; j < P; j++) {
            double s1 = 0.0, s2 = 0.0;
            for (int k = 0; k < P; k++) { s1 = fma(G[i][k], At[k][j], s1); s2 = fma(G[i][k], Ct[k][j], s2); }
            a.A[i][j] = s1;
            GC[i][j] = s2;
        }
    }
    for (int i = 0; i < P; i++)
        for (int j = 0; j < P; j++) {
            double s = 0.0;
            for (int k = 0; k < P; k++) s = fma(GC[i][k], G[j][k], s);
            a.C[i][j] = s + (o.Q[i][j] - o.K[i] * o.Qc[j]);
        }
}

// Minv = (I + C J)^{-1} by Gauss-Jordan with partial pivoting; returns false if singular
template <int P>
__device__ bool inv_I_plus_CJ(const double C[P][P], const double J[P][P], double Minv[P][P]) {
    double W[P][2 * P];
    for (int i = 0; i < P; i++)
        for (int j = 0; j < P; j++) {
            double s = (i == j) ? 1.0 : 0.0;
            for (int k = 0; k < P; k++) s = fma(C[i][k], J[k][j], s);
            W[i][j] = s;
            W[i][P + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int col = 0; col < P; col++) {
        int piv = col;
        double best = fabs(W[col][col]);
        for (int r = col + 1; r < P; r++)
            if (fabs(W[r][col]) > best) { best = fabs(W[r][col]); piv = r; }
        if (!(best > 0.0)) return false;
        if (piv != col)
            for (int j = 0; j < 2 * P; j++) { double t = W[col][j]; W[col][j] = W[piv][j]; W[piv][j] = t; }
        double ip = 1.0 / W[col][col];
        for (int j = 0; j < 2 * P; j++) W[col][j] *= ip;
        for (int r = 0; r < P; r++)
            if (r != col) {
                double f = W[r][col];
                for (int j = 0; j < 2 * P; j++) W[r][j] = fma(-f, W[col][j], W[r][j]);
            }
    }
    for (int i = 0; i < P; i++)
        for (int j = 0; j < P; j++) Minv[i][j] = W[i][P + j];
    return true;
}

// out = i (x) j (general operator).  out may alias neither input.
template <int P>
__device__ void combine(const ScanElem<P>& i, const ScanElem<P>& j, ScanElem<P>& out) {
    double M[P][P];
    inv_I_plus_CJ<P>(i.C, j.J, M);
    double AM[P][P];  // A_j M
    for (int r = 0; r < P; r++)
        for (int c = 0; c < P; c++) {
            double s = 0.0;
            for (int k = 0; k < P; k++) s = fma(j.A[r][k], M[k][c], s);
            AM[r][c] = s;
        }
    double v[P], u[P];
    for (int r = 0; r < P; r++) {
        double s = i.b[r], t = j.eta[r];
        for (int k = 0; k < P; k++) { s = fma(i.C[r][k], j.eta[k], s); t = fma(-j.J[r][k], i.b[k], t); }
        v[r] = s;   // b_i + C_i eta_j
        u[r] = t;   // eta_j - J_j b_i
    }
    double AMC[P][P], MtJ[P][P], Mtu[P];
    for (int r = 0; r < P; r++) {
        double s = j.b[r], t = 0.0;
        for (int k = 0; k < P; k++) { s = fma(AM[r][k], v[k], s); t = fma(M[k][r], u[k], t); }
        out.b[r] = s;
        Mtu[r] = t;
        for (int c = 0; c < P; c++) {
            double s1 = 0.0, s2 = 0.0, s3 = 0.0;
            for (int k = 0; k < P; k++) {
                s1 = fma(AM[r][k], i.A[k][c], s1);
                s2 = fma(AM[r][k], i.C[k][c], s2);
                s3 = fma(M[k][r], j.J[k][c], s3);
            }
            out.A[r][c] = s1;
            AMC[r][c] = s2;
            MtJ[r][c] = s3;
        }
    }
    double MtJA[P][P];
    for (int r = 0; r < P; r++)
        for (int c = 0; c < P; c++) {
            double s1 = j.C[r][c], s2 = 0.0;
            for (int k = 0; k < P; k++) { s1 = fma(AMC[r][k], j.A[c][k], s1); s2 = fma(MtJ[r][k], i.A[k][c], s2); }
            out.C[r][c] = s1;
            MtJA[r][c] = s2;
        }
    for (int r = 0; r < P; r++) {
        double s = i.eta[r];
        for (int k = 0; k < P; k++) s = fma(i.A[k][r], Mtu[k], s);
        out.eta[r] = s;
        for (int c = 0; c < P; c++) {
            double t = i.J[r][c];
            for (int k = 0; k < P; k++) t = fma(i.A[k][r], MtJA[k][c], t);
            out.J[r][c] = t;
        }
    }
}

// filtered state (b, C) pushed through aggregate j
template <int P>
__device__ void apply_elem(FiltState<P>& f, const ScanElem<P>& j) {
    double M[P][P];
    inv_I_plus_CJ<P>(f.C, j.J, M);
    double AM[P][P], v[P];
    for (int r = 0; r < P; r++) {
        double s = f.b[r];
        for (int k = 0; k < P; k++) s = fma(f.C[r][k], j.eta[k], s);
        v[r] = s;
        for (int c = 0; c < P; c++) {
            double t = 0.0;
            for (int k = 0; k < P; k++) t = fma(j.A[r][k], M[k][c], t);
            AM[r][c] = t;
        }
    }
    double AMC[P][P], nb[P];
    for (int r = 0; r < P; r++) {
        double s = j.b[r];
        for (int k = 0; k < P; k++) s = fma(AM[r][k], v[k], s);
        nb[r] = s;
        for (int c = 0; c < P; c++) {
            double t = 0.0;
            for (int k = 0; k < P; k++) t = fma(AM[r][k], f.C[k][c], t);
            AMC[r][c] = t;
        }
    }
    for (int r = 0; r < P; r++) {
        f.b[r] = nb[r];
        for (int c = 0; c < P; c++) {
            double t = j.C[r][c];
            for (int k = 0; k < P; k++) t = fma(AMC[r][k], j.A[c][k], t);
            f.C[r][c] = t;
        }
    }
}

constexpr int SCAN_BLOCK = 128;

// pass 1
template <int P>
__global__ void __launch_bounds__(SCAN_BLOCK)
scan_reduce_kernel(SeriesView sv, const ScanParams<P>* __restrict__ spp, int chunk, int nchunks,
                   ScanElem<P>* __restrict__ E) {
    // blockIdx.y = theta row: every per-row array is offset by it
    MathTab tb;
    tb.load();
    spp += blockIdx.y;
    E += (size_t)blockIdx.y * nchunks;
    const int m = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    if (m >= nchunks || spp->status != TT_OK) return;
    const ScanParams<P>& sp = *spp;
    const int lo = m * chunk, hi = min(sv.ny, lo + chunk);
    ScanElem<P> acc;
    StepOps<P> o;
    const double mu = sp.prm.mu, scale = sp.prm.scale;
    if (lo == 0) {
        first_elem<P>(sp, sv.y[0] - mu, scale * sv.e2_0, acc);
    } else {
        o.build(sp, tb, sv.dt[lo - 1], scale * sv.e2n[lo - 1]);
        single_elem<P>(sp, o, sv.y[lo] - mu, acc);
    }
    for (int k = lo + 1; k < hi; k++) {
        o.build(sp, tb, sv.dt[k - 1], scale * sv.e2n[k - 1]);
        fold_step<P>(o, sv.y[k] - mu, acc);
    }
    E[m] = acc;
}

// pass 2, hierarchical (round 1 ran it in ONE block: 61 % of the whole evaluation at ny = 10^6):
//   2a  scan_block_kernel    every block of 256 aggregates does a Kogge-Stone inclusive scan (8 combine levels, operands
//                            exchanged through an L2-resident double buffer) -> within-block prefixes I[m], block totals;
//   2b  scan_totals_kernel   one block scans the <= 256 block totals the same way -> filtered state at every block start;
//   2c  scan_apply_kernel    every aggregate pushes its block's start state through its within-block prefix: the
//                            filtered state in front of chunk m + 1.
// Sequential depth: 8 + 8 + 1 combines instead of ~70.
constexpr int SCAN_TILE = 256;

template <int P>
__global__ void __launch_bounds__(SCAN_TILE)
scan_block_kernel(const ScanParams<P>* __restrict__ spp, const ScanElem<P>* __restrict__ E, int M, int nblk,
                  ScanElem<P>* __restrict__ X /* 2 M per row */, ScanElem<P>* __restrict__ I /* M per row */,
                  ScanElem<P>* __restrict__ B /* nblk per row */) {
    const int row = blockIdx.y;
    if (spp[row].status != TT_OK) return;
    E += (size_t)row * M;
    X += (size_t)row * 2 * M;
    I += (size_t)row * M;
    B += (size_t)row * nblk;
    const int t = threadIdx.x, m = blockIdx.x * SCAN_TILE + t;
    const bool have = m < M;
    ScanElem<P> acc;
    if (have) acc = E[m];
    const ScanElem<P>* prev = E;
    int lvl = 0;
    for (int off = 1; off < SCAN_TILE; off <<= 1, lvl++) {
        ScanElem<P>* next = X + (size_t)(lvl & 1) * M;
        if (have) {
            if (t >= off) {
                ScanElem<P> out;
                combine<P>(prev[m - off], acc, out);
                acc = out;
            }
            next[m] = acc;
        }
        prev = next;
        __threadfence_block();
        __syncthreads();
    }
    if (have) {
        I[m] = acc;
        if (t == SCAN_TILE - 1 || m == M - 1) B[blockIdx.x] = acc;
    }
}

template <int P>
__global__ void __launch_bounds__(SCAN_TILE)
scan_totals_kernel(const ScanParams<P>* __restrict__ spp, const ScanElem<P>* __restrict__ B, int nblk,
                   ScanElem<P>* __restrict__ BX /* 2 nblk per row */, FiltState<P>* __restrict__ S /* nblk per row */) {
    const int row = blockIdx.x;
    if (spp[row].status != TT_OK) return;
    B += (size_t)row * nblk;
    BX += (size_t)row * 2 * nblk;
    S += (size_t)row * nblk;
    const int t = threadIdx.x;
    const bool have = t < nblk;
    ScanElem<P> acc;
    if (have) acc = B[t];
    const ScanElem<P>* prev = B;
    int lvl = 0;
    for (int off = 1; off < nblk; off <<= 1, lvl++) {
        ScanElem<P>* next = BX + (size_t)(lvl & 1) * nblk;
        if (have) {
            if (t >= off) {
                ScanElem<P> out;
                combine<P>(prev[t - off], acc, out);
                acc = out;
            }
            next[t] = acc;
        }
        prev = next;
        __threadfence_block();
        __syncthreads();
    }
    // state at the START of block t + 1 = (b, C) of the inclusive prefix over blocks 0..t
    if (have && t + 1 < nblk) {
        FiltState<P> f;
        for (int i = 0; i < P; i++) { f.b[i] = acc.b[i]; for (int j = 0; j < P; j++) f.C[i][j] = acc.C[i][j]; }
        S[t + 1] = f;
    }
}

template <int P>
__global__ void __launch_bounds__(SCAN_TILE)
scan_apply_kernel(const ScanParams<P>* __restrict__ spp, const ScanElem<P>* __restrict__ I, const FiltState<P>* __restrict__ S,
                  int M, int nblk, FiltState<P>* __restrict__ F /* M per row */) {
    const int row = blockIdx.y;
    if (spp[row].status != TT_OK) return;
    I += (size_t)row * M;
    S += (size_t)row * nblk;
    F += (size_t)row * M;
    const int m = blockIdx.x * SCAN_TILE + threadIdx.x;
    if (m + 1 >= M) return;  // F[m + 1] is the state in front of chunk m + 1; chunk 0 starts from Reset()
    FiltState<P> f;
    if (blockIdx.x == 0) {
        // no predecessor block: the prefix itself carries the filtered state (its first element is the prior update)
        for (int i = 0; i < P; i++) { f.b[i] = I[m].b[i]; for (int j = 0; j < P; j++) f.C[i][j] = I[m].C[i][j]; }
    } else {
        f = S[blockIdx.x];
        apply_elem<P>(f, I[m]);
    }
    F[m + 1] = f;
}

// What pass 3 can write per point besides the log-likelihood terms: the one-step predictive mean / variance
// (KalmanFilter<>::mean, var: kfilter.hpp:31-32) and the whole predicted state (z, D) -- the latter is what Predict
// resumes from, so that no query re-runs the filter over the points in front of it.
template <int P>
struct ScanEmit {
    double* mean;   // [ny] or nullptr
    double* var;    // [ny] or nullptr
    double* state;  // [ny][STATE_DOUBLES] or nullptr: z[P], D[NT] of the state predicted at point i (before conditioning on it)
    static constexpr int STATE_DOUBLES = P + P * (P + 1) / 2;
};

// pass 3
template <int P, bool EMIT>
__global__ void __launch_bounds__(SCAN_BLOCK)
scan_filter_kernel(SeriesView sv, const ScanParams<P>* __restrict__ spp, int chunk, int nchunks,
                   const FiltState<P>* __restrict__ F, double* __restrict__ LL, ScanEmit<P> em) {
    MathTab tb;
    tb.load();
    spp += blockIdx.y;
    F += (size_t)blockIdx.y * nchunks;
    LL += (size_t)blockIdx.y * nchunks;
    const int m = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    if (m >= nchunks || spp->status != TT_OK) return;
    const RealParams<P> prm = spp->prm;
    const int lo = m * chunk, hi = min(sv.ny, lo + chunk);
    KalmanReal<P> kf;
    LogLikAcc acc;
    acc.init();
    if (m == 0) {
        kf.reset(prm, sv.e2_0);
    } else {
        const FiltState<P>& f = F[m];
#pragma unroll
        for (int i = 0; i < P; i++) {
            kf.z[i] = f.b[i];
#pragma unroll
            for (int j = i; j < P; j++) kf.D[KalmanReal<P>::idx(i, j)] = f.C[i][j] - spp->V[i][j];
        }
        kf.template predict_observe<false>(prm, tb, sv.dt[lo - 1], sv.e2n[lo - 1]);
    }
    const int len = hi - lo;
    const SeriesPtr src{sv.dt + lo, sv.y + lo, sv.e2n + lo};
    if (EMIT) {
        // explicit per-point loop (row 0 only: emission is a single-model operation)
        double ll = 0.0;
        for (int i = 0; i < len; i++) {
            const int gi = lo + i;
            if (em.mean) em.mean[gi] = kf.mean;
            if (em.var) em.var[gi] = kf.var;
            if (em.state) {
                double* st = em.state + (size_t)gi * ScanEmit<P>::STATE_DOUBLES;
#pragma unroll
                for (int k = 0; k < P; k++) st[k] = kf.z[k];
#pragma unroll
                for (int k = 0; k < KalmanReal<P>::NT; k++) st[P + k] = kf.D[k];
            }
            const double innov = (src.y[i] - prm.mu) - kf.mean;
            ll += -0.5 * log(kf.var) - 0.5 * innov * innov / kf.var;
            if (gi + 1 < sv.ny) kf.template advance<false>(prm, tb, innov, 1.0 / kf.var, src.dt[i], src.e[i]);
        }
        LL[m] = ll;
        return;
    }
    // every point of the chunk is scored; the transition out of the last one belongs to the next chunk
    const KalmanReal<P> kf0 = kf;
    filter_span_impl<P, false, true>(kf, acc, prm, tb, src, len, len - 1);
    double ll = acc.value();
    if (acc.bad()) {  // a variance outside the normal range: redo the chunk with one log() per point
        kf = kf0;
        ll = 0.0;
        for (int i = 0; i < len; i++) {
            const double innov = (src.y[i] - prm.mu) - kf.mean;
            ll += -0.5 * log(kf.var) - 0.5 * innov * innov / kf.var;
            if (i + 1 < len) kf.template advance<false>(prm, tb, innov, 1.0 / kf.var, src.dt[i], src.e[i]);
        }
    }
    LL[m] = ll;
}

// pass 4: deterministic fixed-order sum by one block
template <int P>
__global__ void __launch_bounds__(256)
scan_sum_kernel(const ScanParams<P>* __restrict__ spp, const double* __restrict__ LL, int M, double* __restrict__ out) {
    __shared__ double sh[256];
    spp += blockIdx.x;  // one block per theta row
    LL += (size_t)blockIdx.x * M;
    out += blockIdx.x;
    if (spp->status != TT_OK) {
        if (threadIdx.x == 0) *out = -INFINITY;
        return;
    }
    double s = 0.0;
    for (int m = threadIdx.x; m < M; m += 256) s += LL[m];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0] + spp->prm.logprior;
}

// Points folded per thread when the caller does not say (chunk <= 0).  Measured (scripts/scan_chunk_probe.py, one
// CARMA(3,1) theta): ny = 1e5 is fastest at 16 (0.13 ms against 0.25 at 128), ny = 1e6 at 32..64 (0.20 against 0.28),
// ny = 4e6 at 128: the two per-chunk passes are sequential in the chunk, so the chunk should be as short as still
// leaves the aggregate scan small -- about 24,000 aggregates over all rows.
static int scan_auto_chunk(size_t ny, size_t nrows) {
    const size_t target = std::max<size_t>(24000 / std::max<size_t>(nrows, 1), 256);
    const size_t c = (ny + target - 1) / target;
    return (int)std::min<size_t>(std::max<size_t>(c, 16), 128);
}

// The scan passes for `nrows` models whose ScanParams are produced by `fill_params(sp)`; with `em` (nrows must be 1)
// pass 3 also writes the per-point predictive mean / variance / state.
template <int P, class FillParams>
static int scan_core(carma_series* s, int nrows, int chunk, cudaStream_t st, double* d_out, const ScanEmit<P>* em,
                     FillParams fill_params, const double* d_y_override = nullptr) {
    SeriesView sv = s->view();
    if (d_y_override) sv.y = d_y_override;   // same times and errors, other values (conditional simulation)
    if (chunk <= 0) chunk = scan_auto_chunk(sv.ny, (size_t)nrows);
    chunk = std::max(chunk, 2);
    // two scan levels of 256 cover 65,536 aggregates: longer series get longer chunks
    chunk = std::max(chunk, (int)((sv.ny + (SCAN_TILE * SCAN_TILE) - 1) / (SCAN_TILE * SCAN_TILE)));
    const int M = (sv.ny + chunk - 1) / chunk;
    const int nblk = (M + SCAN_TILE - 1) / SCAN_TILE;
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t b_sp = align((size_t)nrows * sizeof(ScanParams<P>));
    const size_t b_E = align((size_t)nrows * M * sizeof(ScanElem<P>));
    const size_t b_X = align((size_t)nrows * 2 * M * sizeof(ScanElem<P>));
    const size_t b_I = align((size_t)nrows * M * sizeof(ScanElem<P>));
    const size_t b_B = align((size_t)nrows * nblk * sizeof(ScanElem<P>));
    const size_t b_BX = align((size_t)nrows * 2 * nblk * sizeof(ScanElem<P>));
    const size_t b_S = align((size_t)nrows * nblk * sizeof(FiltState<P>));
    const size_t b_F = align((size_t)nrows * M * sizeof(FiltState<P>));
    const size_t b_LL = align((size_t)nrows * M * sizeof(double));
    if (!s->scratch_misc.reserve(b_sp + b_E + b_X + b_I + b_B + b_BX + b_S + b_F + b_LL + 256)) return CARMA_ERR_CUDA;
    char* base = (char*)s->scratch_misc.p;
    ScanParams<P>* sp = (ScanParams<P>*)base; base += b_sp;
    ScanElem<P>* E = (ScanElem<P>*)base; base += b_E;
    ScanElem<P>* X = (ScanElem<P>*)base; base += b_X;
    ScanElem<P>* I = (ScanElem<P>*)base; base += b_I;
    ScanElem<P>* B = (ScanElem<P>*)base; base += b_B;
    ScanElem<P>* BX = (ScanElem<P>*)base; base += b_BX;
    FiltState<P>* S = (FiltState<P>*)base; base += b_S;
    FiltState<P>* F = (FiltState<P>*)base; base += b_F;
    double* LL = (double*)base;
    fill_params(sp);
    dim3 grid((unsigned)((M + SCAN_BLOCK - 1) / SCAN_BLOCK), (unsigned)nrows);
    dim3 tgrid((unsigned)nblk, (unsigned)nrows);
    scan_reduce_kernel<P><<<grid, SCAN_BLOCK, 0, st>>>(sv, sp, chunk, M, E);
    scan_block_kernel<P><<<tgrid, SCAN_TILE, 0, st>>>(sp, E, M, nblk, X, I, B);
    scan_totals_kernel<P><<<nrows, SCAN_TILE, 0, st>>>(sp, B, nblk, BX, S);
    scan_apply_kernel<P><<<tgrid, SCAN_TILE, 0, st>>>(sp, I, S, M, nblk, F);
    if (em) scan_filter_kernel<P, true><<<grid, SCAN_BLOCK, 0, st>>>(sv, sp, chunk, M, F, LL, *em);
    else scan_filter_kernel<P, false><<<grid, SCAN_BLOCK, 0, st>>>(sv, sp, chunk, M, F, LL, ScanEmit<P>{nullptr, nullptr, nullptr});
    scan_sum_kernel<P><<<nrows, 256, 0, st>>>(sp, LL, M, d_out);
    return cuda_ok(cudaGetLastError(), "scan kernels launch") ? CARMA_OK : CARMA_ERR_CUDA;
}

// nrows theta rows on one series: 7 launches in total (grid.y / one block per row)
template <int P>
static int scan_rows(carma_series* s, int kind, int q, unsigned flags, const carma_prior_t& prior, const double* d_theta,
                     double* d_out, int nrows, int chunk, cudaStream_t st) {
    const int d = model_dim(kind, P, q);
    const double dt_max = s->dt_max;
    return scan_core<P>(s, nrows, chunk, st, d_out, nullptr, [&](ScanParams<P>* sp) {
        scan_params_kernel<P><<<(nrows + 31) / 32, 32, 0, st>>>(kind, q, d, flags, prior, dt_max, d_theta, sp, nrows);
    });
}

// One explicit model: the time-parallel form of KalmanFilterp::Filter() with its per-point outputs.
// d_mean / d_var: [ny] device arrays or nullptr; d_state: [ny][P + P(P+1)/2] or nullptr; d_loglik: one double.
template <int P>
static int scan_explicit_p(carma_series* s, const ExplicitModel& ex, double* d_mean, double* d_var, double* d_state,
                           double* d_loglik, cudaStream_t st, const double* d_y_override) {
    ScanEmit<P> em{d_mean, d_var, d_state};
    const double dt_max = s->dt_max;
    return scan_core<P>(s, 1, 0, st, d_loglik, &em, [&](ScanParams<P>* sp) {
        scan_params_explicit_kernel<P><<<1, 32, 0, st>>>(ex, dt_max, sp);
    }, d_y_override);
}

int scan_explicit(carma_series* s, int p, const ExplicitModel& ex, double* d_mean, double* d_var, double* d_state,
                  double* d_loglik, cudaStream_t st, const double* d_y_override) {
    switch (p) {
        case 1: return scan_explicit_p<1>(s, ex, d_mean, d_var, d_state, d_loglik, st, d_y_override);
        case 2: return scan_explicit_p<2>(s, ex, d_mean, d_var, d_state, d_loglik, st, d_y_override);
        case 3: return scan_explicit_p<3>(s, ex, d_mean, d_var, d_state, d_loglik, st, d_y_override);
        case 4: return scan_explicit_p<4>(s, ex, d_mean, d_var, d_state, d_loglik, st, d_y_override);
        case 5: return scan_explicit_p<5>(s, ex, d_mean, d_var, d_state, d_loglik, st, d_y_override);
        case 6: return scan_explicit_p<6>(s, ex, d_mean, d_var, d_state, d_loglik, st, d_y_override);
        case 7: return scan_explicit_p<7>(s, ex, d_mean, d_var, d_state, d_loglik, st, d_y_override);
        default: return CARMA_ERR_ARG;
    }
}

}  // namespace carma

using namespace carma;

extern "C" {

int carma_loglik_scan_dev(carma_series_t s, int kind, int p, int q, const carma_prior_t* prior, size_t n,
                          const double* d_theta, double* d_logpost, unsigned flags, int chunk, void* stream) {
    if (!s || !prior || (!d_theta && n) || (!d_logpost && n)) { set_error("carma_loglik_scan_dev: null argument"); return CARMA_ERR_ARG; }
    if (kind < CARMA_KIND_CAR1 || kind > CARMA_KIND_ZCARMA || p < 1 || p > MAX_P || (kind == CARMA_KIND_CAR1 && p != 1) ||
        (kind == CARMA_KIND_CARMA && !(q >= 0 && q < p))) {
        set_error("carma_loglik_scan_dev: invalid (kind,p,q)");
        return CARMA_ERR_ARG;
    }
    if (!cuda_ok(cudaSetDevice(s->device), "cudaSetDevice")) return CARMA_ERR_CUDA;
    const size_t d = (size_t)model_dim(kind, p, q);
    cudaStream_t st = (cudaStream_t)stream;
    // rows are processed in groups so that the per-row scratch stays below ~1 GiB (and grid.y <= 65535)
    if (chunk <= 0) chunk = scan_auto_chunk(s->ny, n);
    const size_t chunk_eff = (size_t)std::max(chunk, 2);
    size_t per_row = ((size_t)s->ny / chunk_eff + 2) * (size_t)(4 * p * p + 3 * p + 1) * sizeof(double) * 5;
    size_t group = std::max<size_t>(1, std::min<size_t>(n, std::min<size_t>(4096, ((size_t)1 << 30) / std::max<size_t>(per_row, 1))));
    for (size_t i0 = 0; i0 < n; i0 += group) {
        const int nr = (int)std::min(group, n - i0);
        const double* th = d_theta + i0 * d;
        double* out = d_logpost + i0;
        int rc;
        switch (p) {
            case 1: rc = scan_rows<1>(s, kind, q, flags, *prior, th, out, nr, chunk, st); break;
            case 2: rc = scan_rows<2>(s, kind, q, flags, *prior, th, out, nr, chunk, st); break;
            case 3: rc = scan_rows<3>(s, kind, q, flags, *prior, th, out, nr, chunk, st); break;
            case 4: rc = scan_rows<4>(s, kind, q, flags, *prior, th, out, nr, chunk, st); break;
            case 5: rc = scan_rows<5>(s, kind, q, flags, *prior, th, out, nr, chunk, st); break;
            case 6: rc = scan_rows<6>(s, kind, q, flags, *prior, th, out, nr, chunk, st); break;
            case 7: rc = scan_rows<7>(s, kind, q, flags, *prior, th, out, nr, chunk, st); break;
            default: rc = CARMA_ERR_ARG;
        }
        if (rc) return rc;
    }
    return CARMA_OK;
}

int carma_loglik_scan(carma_series_t s, int kind, int p, int q, const carma_prior_t* prior, size_t n,
                      const double* theta, double* logpost, unsigned flags, int chunk) {
    if (!s || !prior || (!theta && n) || (!logpost && n)) { set_error("carma_loglik_scan: null argument"); return CARMA_ERR_ARG; }
    if (n == 0) return CARMA_OK;
    if (!cuda_ok(cudaSetDevice(s->device), "cudaSetDevice")) return CARMA_ERR_CUDA;
    if (p < 1 || p > MAX_P) { set_error("carma_loglik_scan: invalid p"); return CARMA_ERR_ARG; }
    size_t d = (size_t)model_dim(kind, p, q);
    if (!s->scratch_in.reserve(n * d * sizeof(double)) || !s->scratch_out.reserve(n * sizeof(double))) return CARMA_ERR_CUDA;
    if (!cuda_ok(cudaMemcpy(s->scratch_in.p, theta, n * d * sizeof(double), cudaMemcpyHostToDevice), "H2D theta")) return CARMA_ERR_CUDA;
    int rc = carma_loglik_scan_dev(s, kind, p, q, prior, n, (const double*)s->scratch_in.p, (double*)s->scratch_out.p, flags, chunk, 0);
    if (rc) return rc;
    if (!cuda_ok(cudaMemcpy(logpost, s->scratch_out.p, n * sizeof(double), cudaMemcpyDeviceToHost), "D2H logpost")) return CARMA_ERR_CUDA;
    return CARMA_OK;
}

}  // extern "C"
