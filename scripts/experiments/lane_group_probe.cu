// micro-benchmark: cycles per Kalman step of ONE chain, lane-group form vs one-thread form (lone warp)
#include <cstdio>
#include <vector>
#include <cmath>
#include "kalman_group.cuh"
using namespace carma;

__global__ void probe(const double* th, const double* sdt, const double* sy, const double* se, int ny, double* out, long long* cyc, int mode, int vary) {
    extern __shared__ double sm[];
    for (int k = threadIdx.x; k < ny; k += blockDim.x) { sm[k] = sdt[k]; sm[ny + k] = sy[k]; sm[2 * ny + k] = se[k]; }
    __syncthreads();
    const double *pdt = sm, *py = sm + ny, *pe = sm + 2 * ny;
    carma_prior_t pr{1e300, 1e300, 0, 0, 1, 50};
    RealParams<5> prm;
    double t[MAX_D];
    for (int j = 0; j < MAX_D; j++) t[j] = j < 11 ? th[j] : 0.0;
    if (vary && mode == 1) { t[3] += 0.01 * threadIdx.x; t[5] -= 0.02 * threadIdx.x; if (vary == 2 && threadIdx.x == 3) { t[3] = log(0.02); t[4] = log(0.9); } }
    if (transform_theta<5>(CARMA_KIND_CARMA, 3, 1u, pr, t, prm) != TT_OK) return;
    long long t0 = clock64();
    double r;
    if (mode == 0) {
        const int l16 = threadIdx.x & 15;
        const UnitPar ua = unit_of<5>(prm, l16 >> 2), ub = unit_of<5>(prm, l16 & 3);
        r = group_filter(ua, ub, prm.v0, prm.scale, prm.mu, pdt, py, pe, 0.01, ny, l16);
    } else {
        KalmanReal<5> kf;
        LogLikAcc acc;
        kf.reset(prm, 0.01);
        acc.init();
        filter_span_any<5, true>(kf, acc, prm, pdt, py, pe, ny, ny - 1);
        r = acc.value();
    }
    long long t1 = clock64();
    out[threadIdx.x] = r;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    const int ny = 1000;
    std::vector<double> dt(ny), y(ny), e(ny);
    unsigned s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 8) / 16777216.0; };
    for (int i = 0; i < ny; i++) { dt[i] = 0.5 + rnd(); y[i] = 17.0 + 2.0 * (rnd() - 0.5); e[i] = 0.01; }
    double th[11] = {2.3, 1.0, 17.0, 0.4593751881490773, -2.0741459390188006, -2.755077074073137, -3.17275822768691,
                     -3.460440300138691, -0.2231435513142097, 1.2809338454620642, 3.912023005428146};
    double *d_th, *d_dt, *d_y, *d_e, *d_out; long long* d_c;
    cudaMalloc(&d_th, sizeof(th)); cudaMalloc(&d_dt, ny * 8); cudaMalloc(&d_y, ny * 8); cudaMalloc(&d_e, ny * 8);
    cudaMalloc(&d_out, 64 * 8); cudaMalloc(&d_c, 8);
    cudaMemcpy(d_th, th, sizeof(th), cudaMemcpyHostToDevice);
    cudaMemcpy(d_dt, dt.data(), ny * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_y, y.data(), ny * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_e, e.data(), ny * 8, cudaMemcpyHostToDevice);
    for (int vary = 0; vary < 3; vary++)
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 2; rep++) probe<<<1, 32, 3 * ny * 8>>>(d_th, d_dt, d_y, d_e, ny, d_out, d_c, mode, vary);
        cudaDeviceSynchronize();
        double out[32]; long long c;
        cudaMemcpy(out, d_out, sizeof(out), cudaMemcpyDeviceToHost);
        cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
        printf("vary=%d %s: loglik %.12f (lane 17: %.12f)  %.1f cycles/step  err=%s\n", vary, mode == 0 ? "lane-group" : "one-thread", out[0], out[17],
               (double)c / ny, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
