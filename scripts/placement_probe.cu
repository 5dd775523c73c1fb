// Where do the warps of a partially filled grid land?  Every block spins long enough for the whole grid to be
// co-resident and records (%smid, %warpid) of each of its warps; the host prints, per grid size, the histogram of
// blocks per SM and of warps per scheduler slot class (%warpid mod 4 = the SM sub-partition on sm_100).
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o placement_probe placement_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(64) probe(int* smid, int* warpid, long long spin) {
    extern __shared__ double pad[];
    unsigned s, w;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(w));
    long long t0 = clock64();
    while (clock64() - t0 < spin) { }
    if ((threadIdx.x & 31) == 0) {
        int k = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
        smid[k] = (int)s;
        warpid[k] = (int)w;
    }
    if (spin < 0) pad[threadIdx.x] = 0.0;
}

int main(int argc, char** argv) {
    int smem = argc > 1 ? atoi(argv[1]) : 34 * 1024;
    int block = argc > 2 ? atoi(argv[2]) : 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int grids[] = {148, 296, 444, 592, 683, 740, 888, 1024};
    for (int g : grids) {
        int nw = g * (block / 32);
        int *d_s, *d_w;
        cudaMalloc(&d_s, nw * sizeof(int));
        cudaMalloc(&d_w, nw * sizeof(int));
        probe<<<g, block, smem>>>(d_s, d_w, 2000000LL);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<int> s(nw), w(nw);
        cudaMemcpy(s.data(), d_s, nw * sizeof(int), cudaMemcpyDeviceToHost);
        cudaMemcpy(w.data(), d_w, nw * sizeof(int), cudaMemcpyDeviceToHost);
        std::map<int, int> per_sm;
        std::map<int, std::vector<int>> slot;   // sm -> warps per (warpid & 3)
        for (int k = 0; k < nw; k++) {
            per_sm[s[k]]++;
            auto& v = slot[s[k]];
            if (v.empty()) v.assign(4, 0);
            v[w[k] & 3]++;
        }
        std::map<int, int> hist_sm, hist_max;
        std::map<std::string, int> pattern;
        for (auto& kv : per_sm) hist_sm[kv.second]++;
        for (auto& kv : slot) {
            int mx = 0;
            char buf[64];
            for (int x : kv.second) mx = x > mx ? x : mx;
            snprintf(buf, sizeof buf, "(%d,%d,%d,%d)", kv.second[0], kv.second[1], kv.second[2], kv.second[3]);
            pattern[buf]++;
            hist_max[mx]++;
        }
        printf("grid %4d x %d threads, smem %d: SMs used %zu; warps per SM -> #SMs:", g, block, smem, per_sm.size());
        for (auto& kv : hist_sm) printf(" %d:%d", kv.first, kv.second);
        printf(" | max warps on one (warpid mod 4) class -> #SMs:");
        for (auto& kv : hist_max) printf(" %d:%d", kv.first, kv.second);
        printf(" | patterns:");
        for (auto& kv : pattern) printf(" %s x%d", kv.first.c_str(), kv.second);
        printf("\n");
        cudaFree(d_s); cudaFree(d_w);
    }
    return 0;
}
