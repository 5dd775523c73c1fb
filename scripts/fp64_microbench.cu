// FP64 pipe micro-benchmarks on B200: dependent-issue latency and throughput vs warps/scheduler and ILP.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_microbench fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_chains(double* out, int iters, double a, double b) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = threadIdx.x * 1e-3 + c;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) x[c] = fma(x[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += x[c];
    if (s == 12345.678) out[0] = s;
}

// mixed: DMUL + DFMA with all-distinct register operands (no loop-invariant operands)
template <int CHAINS>
__global__ void dfma_distinct(double* out, int iters) {
    double x[CHAINS], y[CHAINS], z[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { x[c] = threadIdx.x * 1e-3 + c; y[c] = 1.0 + 1e-9 * c; z[c] = 1e-9 * c; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) { x[c] = fma(x[c], y[c], z[c]); y[c] = fma(y[c], z[c], x[c]); z[c] = fma(z[c], x[c], y[c]); }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) s += x[c] + y[c] + z[c];
    if (s == 12345.678) out[0] = s;
}

template <int CHAINS>
void run(const char* name, int warps_per_sm, int nsm, double clock_ghz, bool distinct) {
    double* d; cudaMalloc(&d, 8);
    int iters = 1 << 14;
    int threads = 32 * warps_per_sm > 1024 ? 1024 : 32 * warps_per_sm;
    int blocks_per_sm = (32 * warps_per_sm + threads - 1) / threads;
    int blocks = nsm * blocks_per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (distinct) dfma_distinct<CHAINS><<<blocks, threads>>>(d, iters);
        else dfma_chains<CHAINS><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fmas = (double)blocks * threads * iters * CHAINS * (distinct ? 3 : 1);
    double per_clk_sm = fmas / (ms * 1e-3) / (clock_ghz * 1e9) / nsm;
    double cyc_per_iter_warp = ms * 1e-3 * clock_ghz * 1e9 / iters;  // cycles per loop iteration (all warps concurrent)
    printf("%-10s warps/SM %3d chains %2d : %6.2f FMA lanes/clk/SM  (%5.1f TFLOP/s)  %7.1f cyc/iter\n", name, warps_per_sm,
           CHAINS * (distinct ? 3 : 1), per_clk_sm, 2 * fmas / (ms * 1e-3) / 1e12, cyc_per_iter_warp);
    cudaFree(d);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount; double ghz = p.clockRate * 1e-6;
    printf("%s SMs %d clock %.3f GHz\n", p.name, nsm, ghz);
    int ws[] = {4, 8, 12, 16, 32, 64};
    for (int w : ws) {
        run<1>("invariant", w, nsm, ghz, false);
        run<2>("invariant", w, nsm, ghz, false);
        run<4>("invariant", w, nsm, ghz, false);
        run<8>("invariant", w, nsm, ghz, false);
        run<16>("invariant", w, nsm, ghz, false);
        run<1>("distinct", w, nsm, ghz, true);
        run<2>("distinct", w, nsm, ghz, true);
        run<4>("distinct", w, nsm, ghz, true);
    }
    return 0;
}
