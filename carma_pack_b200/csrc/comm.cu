// comm.cu -- K6: the only collective of the system, behind the C ABI.  Independent units (theta rows, PT ensembles,
// light curves, (p,q,start) fits) are partitioned over one process per GPU with no data-path collective; at the end
// each rank contributes a small summary (per model: -loglik, AICc, theta-hat; per survey shard: a few moments) and
// every rank receives all of them: ONE ncclAllGather of a few kB over NVLink / NVSwitch.  It replaces the pickled
// results that the reference's multiprocessing.Pool sends back through pipes (src/carmcmc/carma_pack.py:111-119).
//
// NCCL is loaded lazily (dlopen "libnccl.so.2"), so libcarma_b200.so has no link-time dependency on it and still
// loads on a box without NCCL; the entry points then return CARMA_ERR_CUDA with an explanatory message.  A host that
// already owns an ncclComm_t (e.g. created next to torch.distributed) passes it straight to carma_gather_summaries;
// a host without one bootstraps with carma_comm_unique_id / carma_comm_init_rank.
#include <cstdlib>
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <string>

#include "series.h"

using namespace carma;

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;
};

NcclApi& api() {
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        // RTLD_NOLOAD first: reuse the NCCL a host framework (torch) already mapped, so both see the same library
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            a.lib = dlopen(name, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
            if (a.lib) break;
        }
        // CARMA_NCCL_LIB: the library to map when none is loaded yet.  A process that imports torch LATER needs torch's
        // bundled NCCL to be the one behind the soname libnccl.so.2 (libtorch_cuda.so binds to newer symbols than an
        // older system NCCL exports); carma_pack_b200/_lib.py points this variable at the bundled copy when there is one.
        if (!a.lib) {
            const char* path = getenv("CARMA_NCCL_LIB");
            if (path && path[0]) a.lib = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
        }
        if (!a.lib)
            for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
                a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
                if (a.lib) break;
            }
        if (!a.lib) { a.why = std::string("NCCL not found: ") + (dlerror() ? dlerror() : "dlopen failed"); return; }
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.lib, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.lib, "ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
        a.CommCount = (decltype(a.CommCount))dlsym(a.lib, "ncclCommCount");
        a.AllGather = (decltype(a.AllGather))dlsym(a.lib, "ncclAllGather");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.lib, "ncclGetErrorString");
        if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.CommCount || !a.AllGather) {
            a.why = "NCCL library lacks a required symbol";
            a.lib = nullptr;
        }
    });
    return a;
}

bool nccl_ok(ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return true;
    NcclApi& a = api();
    set_error(std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(r) : "NCCL error"));
    return false;
}

bool have_nccl() {
    NcclApi& a = api();
    if (a.lib) return true;
    set_error(a.why);
    return false;
}

}  // namespace

extern "C" {

int carma_comm_unique_id(char id[CARMA_COMM_ID_BYTES]) {
    if (!id) return CARMA_ERR_ARG;
    if (!have_nccl()) return CARMA_ERR_CUDA;
    static_assert(CARMA_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "id size");
    ncclUniqueId u;
    if (!nccl_ok(api().GetUniqueId(&u), "ncclGetUniqueId")) return CARMA_ERR_CUDA;
    memcpy(id, u.internal, NCCL_UNIQUE_ID_BYTES);
    return CARMA_OK;
}

int carma_comm_init_rank(int nranks, int rank, const char id[CARMA_COMM_ID_BYTES], int device, void** comm) {
    if (!id || !comm || nranks < 1 || rank < 0 || rank >= nranks) { set_error("carma_comm_init_rank: bad argument"); return CARMA_ERR_ARG; }
    if (!have_nccl()) return CARMA_ERR_CUDA;
    if (!cuda_ok(cudaSetDevice(device), "cudaSetDevice")) return CARMA_ERR_CUDA;
    ncclUniqueId u;
    memcpy(u.internal, id, NCCL_UNIQUE_ID_BYTES);
    ncclComm_t c = nullptr;
    if (!nccl_ok(api().CommInitRank(&c, nranks, u, rank), "ncclCommInitRank")) return CARMA_ERR_CUDA;
    *comm = (void*)c;
    return CARMA_OK;
}

int carma_comm_destroy(void* comm) {
    if (!comm) return CARMA_OK;
    if (!have_nccl()) return CARMA_ERR_CUDA;
    return nccl_ok(api().CommDestroy((ncclComm_t)comm), "ncclCommDestroy") ? CARMA_OK : CARMA_ERR_CUDA;
}

int carma_gather_summaries(void* nccl_comm, const double* local, size_t count, double* all, void* stream) {
    if (!nccl_comm || !local || !all || count == 0) { set_error("carma_gather_summaries: bad argument"); return CARMA_ERR_ARG; }
    if (!have_nccl()) return CARMA_ERR_CUDA;
    ncclComm_t c = (ncclComm_t)nccl_comm;
    int nranks = 0;
    if (!nccl_ok(api().CommCount(c, &nranks), "ncclCommCount")) return CARMA_ERR_CUDA;
    cudaStream_t st = (cudaStream_t)stream;
    double *d_in = nullptr, *d_out = nullptr;
    bool ok = cuda_ok(cudaMalloc((void**)&d_in, count * sizeof(double)), "cudaMalloc(gather in)") &&
              cuda_ok(cudaMalloc((void**)&d_out, (size_t)nranks * count * sizeof(double)), "cudaMalloc(gather out)") &&
              cuda_ok(cudaMemcpyAsync(d_in, local, count * sizeof(double), cudaMemcpyHostToDevice, st), "H2D summary");
    if (ok) ok = nccl_ok(api().AllGather(d_in, d_out, count, ncclDouble, c, st), "ncclAllGather");
    if (ok) ok = cuda_ok(cudaMemcpyAsync(all, d_out, (size_t)nranks * count * sizeof(double), cudaMemcpyDeviceToHost, st), "D2H summaries") &&
                 cuda_ok(cudaStreamSynchronize(st), "gather sync");
    if (d_in) cudaFree(d_in);
    if (d_out) cudaFree(d_out);
    return ok ? CARMA_OK : CARMA_ERR_CUDA;
}

int carma_gather_summaries_dev(void* nccl_comm, const double* d_local, size_t count, double* d_all, void* stream) {
    if (!nccl_comm || !d_local || !d_all || count == 0) { set_error("carma_gather_summaries_dev: bad argument"); return CARMA_ERR_ARG; }
    if (!have_nccl()) return CARMA_ERR_CUDA;
    return nccl_ok(api().AllGather(d_local, d_all, count, ncclDouble, (ncclComm_t)nccl_comm, (cudaStream_t)stream), "ncclAllGather")
               ? CARMA_OK : CARMA_ERR_CUDA;
}

}  // extern "C"
