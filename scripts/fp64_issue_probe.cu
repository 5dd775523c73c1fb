// FP64 issue-rate probe on B200: how many cycles does one warp-wide FP64 instruction occupy the pipe,
// as a function of how many DISTINCT register operands it reads, and does interleaved integer / LDS work
// steal from it?  Complements fp64_microbench.cu (latency + the two extreme operand patterns).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_issue_probe fp64_issue_probe.cu
// Inspect the SASS of the loops (cuobjdump -sass) to see which operands got the .reuse flag.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int C = 12;  // independent chains per thread (latency 9.4 cycles / 2 cycles per issue -> >= 5 needed)

enum Kind { K_DADD2 = 0, K_DMUL2, K_DFMA3, K_DFMA_SHARE1, K_DFMA_SHARE2, K_DFMA_CONST, K_DFMA3_INT, K_DFMA3_LDS,
            K_DFMA_SHARE1_INT, K_NKINDS };
static const char* kNames[] = {"dadd x+=y[c]        (2 regs)", "dmul x*=y[c]        (2 regs)",
                               "dfma x=x*y[c]+z[c]  (3 regs)", "dfma x=x*Y+z[c]     (3 regs, Y shared)",
                               "dfma x=x*Y+Z        (3 regs, Y,Z shared)", "dfma x=x*K+z[c]     (2 regs + imm)",
                               "dfma 3 regs + 1 LOP3 each", "dfma 3 regs + 1 LDS.64 per 4",
                               "dfma Y shared + 1 LOP3 each"};

template <int KIND>
__global__ void probe(double* out, const double* in, int iters) {
    __shared__ double sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = in[threadIdx.x];
    __syncthreads();
    double x[C], y[C], z[C];
#pragma unroll
    for (int c = 0; c < C; c++) { x[c] = in[c] + threadIdx.x * 1e-6; y[c] = in[16 + c] + threadIdx.x * 1e-12; z[c] = in[32 + c] + threadIdx.x * 1e-13; }
    const double Y = in[60] + threadIdx.x * 1e-12, Z = in[61] + threadIdx.x * 1e-13;
    unsigned u = threadIdx.x, v = (unsigned)in[62] + threadIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int c = 0; c < C; c++) {
            if (KIND == K_DADD2) x[c] = x[c] + y[c];
            if (KIND == K_DMUL2) x[c] = x[c] * y[c];
            if (KIND == K_DFMA3 || KIND == K_DFMA3_INT || KIND == K_DFMA3_LDS) x[c] = fma(x[c], y[c], z[c]);
            if (KIND == K_DFMA_SHARE1 || KIND == K_DFMA_SHARE1_INT) x[c] = fma(x[c], Y, z[c]);
            if (KIND == K_DFMA_SHARE2) x[c] = fma(x[c], Y, Z);
            if (KIND == K_DFMA_CONST) x[c] = fma(x[c], 1.0000001, z[c]);
            if (KIND == K_DFMA3_INT || KIND == K_DFMA_SHARE1_INT) u = (u ^ v) + (u >> 3);
            if (KIND == K_DFMA3_LDS && (c & 3) == 0) z[c] = sm[(i + c) & 63];
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < C; c++) s += x[c];
    if (s == 12345.678 || u == 0x12345u) out[0] = s + u;
}

template <int KIND>
void run(int warps_per_sm, int nsm, double ghz, double* d_out, double* d_in) {
    int iters = 1 << 13;
    int threads = 32 * warps_per_sm > 1024 ? 1024 : 32 * warps_per_sm;
    int blocks = nsm * ((32 * warps_per_sm + threads - 1) / threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        probe<KIND><<<blocks, threads>>>(d_out, d_in, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    double inst_per_smsp = (double)warps_per_sm / 4.0 * iters * C;  // FP64 warp-instructions per scheduler
    double cycles = ms * 1e-3 * ghz * 1e9;
    printf("%-44s warps/SMSP %5.2f : %5.2f cycles per FP64 warp-instruction (%5.1f TFLOP/s-equivalent as FMA)\n", kNames[KIND],
           warps_per_sm / 4.0, cycles / inst_per_smsp, 2.0 * 32 * inst_per_smsp * 4 * nsm / (ms * 1e-3) / 1e12);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    double ghz = p.clockRate * 1e-6;
    printf("%s SMs %d clock %.3f GHz, %d chains per thread\n", p.name, nsm, ghz, C);
    double h_in[64];
    for (int i = 0; i < 64; i++) h_in[i] = 1.0 + 1e-9 * i;
    h_in[62] = 12345.0;
    double *d_in, *d_out;
    cudaMalloc(&d_in, sizeof(h_in));
    cudaMalloc(&d_out, 8);
    cudaMemcpy(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
    int ws[] = {4, 8, 14, 16, 32};
    for (int w : ws) {
        run<K_DADD2>(w, nsm, ghz, d_out, d_in);
        run<K_DMUL2>(w, nsm, ghz, d_out, d_in);
        run<K_DFMA3>(w, nsm, ghz, d_out, d_in);
        run<K_DFMA_SHARE1>(w, nsm, ghz, d_out, d_in);
        run<K_DFMA_SHARE2>(w, nsm, ghz, d_out, d_in);
        run<K_DFMA_CONST>(w, nsm, ghz, d_out, d_in);
        run<K_DFMA3_INT>(w, nsm, ghz, d_out, d_in);
        run<K_DFMA3_LDS>(w, nsm, ghz, d_out, d_in);
        run<K_DFMA_SHARE1_INT>(w, nsm, ghz, d_out, d_in);
        printf("\n");
    }
    return 0;
}
