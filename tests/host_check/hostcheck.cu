// hostcheck.cu -- TEST INFRASTRUCTURE (never part of the product library).
// The device headers of the hot path (theta_transform.cuh, kalman_real.cuh, fast_math.cuh) are written as
// __host__ __device__ code; this file instantiates their HOST side behind a tiny C ABI so that the CPU test suite
// (tests/test_host_kernel_math.py, no GPU needed) can compare the exact recursion the kernels run -- real-half
// state, sum/difference basis for real root pairs, pre-scaled range reductions, mantissa-product log sum -- with the
// oracle.  Build:  nvcc -O2 -std=c++17 -Xcompiler -fPIC -shared -o libhostcheck.so hostcheck.cu   (host code only)
#include <cmath>
#include <cstddef>

#include "../../carma_pack_b200/csrc/kalman_real.cuh"

using namespace carma;

template <int P>
static double one(int kind, int q, unsigned flags, const carma_prior_t& pr, const double* th, const double* dt,
                  const double* y, const double* e2n, int ny, double e2_0, double dt_max, int force_generic, int force_slow) {
    RealParams<P> prm;
    if (transform_theta<P>(kind, q, flags, pr, th, dt_max, prm) != TT_OK) return -INFINITY;
    MathTab tb;
    tb.load();
    SeriesPtr src{dt, y, e2n};
    if (force_slow) return loglik_exact_slow<P>(prm, tb, src, ny, e2_0) + prm.logprior;
    KalmanReal<P> kf;
    LogLikAcc acc;
    kf.reset(prm, e2_0);
    acc.init();
    if (force_generic == 2) filter_span_any_pipelined<P>(kf, acc, prm, tb, src, ny, ny - 1);
    else if (force_generic) filter_span_impl<P, false, false>(kf, acc, prm, tb, src, ny, ny - 1);
    else filter_span_any<P, false>(kf, acc, prm, tb, src, ny, ny - 1);
    if (acc.bad()) return loglik_exact_slow<P>(prm, tb, src, ny, e2_0) + prm.logprior;
    return acc.value() + prm.logprior;
}

extern "C" int hostcheck_loglik(int kind, int p, int q, unsigned flags, const carma_prior_t* prior, const double* t,
                                const double* y, const double* yerr, size_t ny, const double* theta, size_t n,
                                double* out, int force_generic, int force_slow) {
    if (p < 1 || p > MAX_P || ny < 2) return 1;
    double* dt = new double[ny];
    double* e2n = new double[ny];
    double dt_max = 0.0;
    for (size_t i = 0; i < ny; i++) {
        dt[i] = (i + 1 < ny) ? t[i + 1] - t[i] : 0.0;
        e2n[i] = (i + 1 < ny) ? yerr[i + 1] * yerr[i + 1] : 0.0;
        if (dt[i] > dt_max) dt_max = dt[i];
    }
    const double e2_0 = yerr[0] * yerr[0];
    const int d = model_dim(kind, p, q);
    for (size_t r = 0; r < n; r++) {
        double th[MAX_D] = {0};
        for (int j = 0; j < d; j++) th[j] = theta[r * d + j];
        double v;
        switch (p) {
            case 1: v = one<1>(kind, q, flags, *prior, th, dt, y, e2n, (int)ny, e2_0, dt_max, force_generic, force_slow); break;
            case 2: v = one<2>(kind, q, flags, *prior, th, dt, y, e2n, (int)ny, e2_0, dt_max, force_generic, force_slow); break;
            case 3: v = one<3>(kind, q, flags, *prior, th, dt, y, e2n, (int)ny, e2_0, dt_max, force_generic, force_slow); break;
            case 4: v = one<4>(kind, q, flags, *prior, th, dt, y, e2n, (int)ny, e2_0, dt_max, force_generic, force_slow); break;
            case 5: v = one<5>(kind, q, flags, *prior, th, dt, y, e2n, (int)ny, e2_0, dt_max, force_generic, force_slow); break;
            case 6: v = one<6>(kind, q, flags, *prior, th, dt, y, e2n, (int)ny, e2_0, dt_max, force_generic, force_slow); break;
            default: v = one<7>(kind, q, flags, *prior, th, dt, y, e2n, (int)ny, e2_0, dt_max, force_generic, force_slow); break;
        }
        out[r] = v;
    }
    delete[] dt;
    delete[] e2n;
    return 0;
}

// the loop's transcendental primitives, element-wise: rate l (table steps per unit time) and dt
extern "C" void hostcheck_fastmath(const double* l, const double* dt, size_t n, double* out_exp, double* out_s_conj,
                                   double* out_c_conj, double* out_s_real, double* out_c_real, double* out_rcp) {
    MathTab tb;
    tb.load();
    for (size_t i = 0; i < n; i++) {
        out_exp[i] = exp_scaled(l[i], dt[i], tb);
        rot_scaled<true>(l[i], dt[i], true, tb, &out_s_conj[i], &out_c_conj[i]);
        double s1, c1;
        rot_scaled<false>(l[i], dt[i], true, tb, &s1, &c1);
        if (s1 != out_s_conj[i] || c1 != out_c_conj[i]) out_s_conj[i] = NAN;  // the two variants must agree bit for bit
        rot_scaled<false>(l[i], dt[i], false, tb, &out_s_real[i], &out_c_real[i]);
        out_rcp[i] = rcp_fast(l[i]);
    }
}
