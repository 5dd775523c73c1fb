// Does a warp-wide FP64 instruction cost less when only part of the warp is active?  One warp per SM sub-partition
// (128-thread blocks, one per SM), 12 independent DFMA chains per thread; the set of active lanes varies.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_lanes_probe scripts/fp64_lanes_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int C = 12;

__global__ void probe(double* out, const double* in, int iters, unsigned lane_mask, long long* cycles) {
    const int lane = threadIdx.x & 31;
    double x[C], y[C], z[C];
#pragma unroll
    for (int c = 0; c < C; c++) { x[c] = in[c] + threadIdx.x * 1e-6; y[c] = in[16 + c] + threadIdx.x * 1e-12; z[c] = in[32 + c] + threadIdx.x * 1e-13; }
    long long t0 = 0, t1 = 0;
    if ((lane_mask >> lane) & 1u) {
        t0 = clock64();
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int c = 0; c < C; c++) x[c] = fma(x[c], y[c], z[c]);
        }
        t1 = clock64();
        double s = 0;
#pragma unroll
        for (int c = 0; c < C; c++) s += x[c];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
        if (blockIdx.x == 0 && threadIdx.x == __ffs(lane_mask) - 1) *cycles = t1 - t0;
    }
}

int main() {
    double *d_in, *d_out, h_in[64];
    long long* d_cyc;
    for (int i = 0; i < 64; i++) h_in[i] = 1.0 + 1e-9 * i;
    cudaMalloc(&d_in, sizeof h_in);
    cudaMalloc(&d_out, 148 * 128 * sizeof(double));
    cudaMalloc(&d_cyc, sizeof(long long));
    cudaMemcpy(d_in, h_in, sizeof h_in, cudaMemcpyHostToDevice);
    const int iters = 20000;
    struct { const char* name; unsigned mask; } cases[] = {
        {"32 lanes", 0xffffffffu}, {"lanes 0-15", 0x0000ffffu}, {"lanes 16-31", 0xffff0000u}, {"even lanes (16)", 0x55555555u},
        {"lanes 0-7", 0x000000ffu}, {"lanes 0-19 (20)", 0x000fffffu}, {"lanes 0-23 (24)", 0x00ffffffu}, {"lane 0", 0x1u}};
    for (int threads : {128, 256}) {
        for (auto& cs : cases) {
            probe<<<148, threads>>>(d_out, d_in, 100, cs.mask, d_cyc);
            cudaDeviceSynchronize();
            probe<<<148, threads>>>(d_out, d_in, iters, cs.mask, d_cyc);
            cudaError_t e = cudaDeviceSynchronize();
            long long cyc = 0;
            cudaMemcpy(&cyc, d_cyc, sizeof cyc, cudaMemcpyDeviceToHost);
            printf("%d warps per sub-partition, %-18s: %.3f cycles per warp DFMA (%s)\n", threads / 128, cs.name,
                   (double)cyc / ((double)iters * C), cudaGetErrorString(e));
        }
    }
    return 0;
}
