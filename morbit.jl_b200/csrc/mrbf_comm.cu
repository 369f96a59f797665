// Multi-GPU plumbing of the C ABI: the final gather of per-instance result rows -- the only collective of the
// path.  Instances shard over GPUs with no data-path collective (examples/large_scale_benchmarks.jl:253 runs them independently);
// when the runs are over, every rank contributes its rows (x, f(x), stop code, #evals, training ids ...) and receives all of them.
//
// NCCL is resolved at run time (dlopen) so that libmorbit_rbf.so has no link-time dependency on it: a single-GPU user never
// needs the library, a multi-GPU host (Julia with NCCL_jll, Python with torch's bundled copy) already has it in the process.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include <cuda_runtime.h>

#include "../../include/morbit_rbf.h"

namespace {

// the slice of nccl.h this file needs (NCCL 2.x ABI: these values and layouts have been stable since 2.0)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt32 = 2, ncclFloat64 = 8 };

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    char err[256] = {0};
};

NcclApi* nccl_api() {
    static NcclApi api;          // function-local static: initialised once, thread-safe (C++11)
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    const char* env = getenv("MRBF_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        api.lib = dlopen(nm, RTLD_NOW | RTLD_NOLOAD);           // the copy the host process already loaded, if any
        if (!api.lib) api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) { snprintf(api.err, sizeof(api.err), "libnccl.so.2 not found (set MRBF_NCCL_LIB): %s", dlerror()); return nullptr; }
    api.GetUniqueId = (ncclResult_t(*)(ncclUniqueId*))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (ncclResult_t(*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(api.lib, "ncclCommInitRank");
    api.CommDestroy = (ncclResult_t(*)(ncclComm_t))dlsym(api.lib, "ncclCommDestroy");
    api.AllGather = (ncclResult_t(*)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t))dlsym(api.lib, "ncclAllGather");
    api.GetErrorString = (const char* (*)(ncclResult_t))dlsym(api.lib, "ncclGetErrorString");
    api.GetVersion = (ncclResult_t(*)(int*))dlsym(api.lib, "ncclGetVersion");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather) {
        snprintf(api.err, sizeof(api.err), "libnccl found but a required symbol is missing");
        api.lib = nullptr;
        return nullptr;
    }
    return &api;
}

}  // namespace

struct mrbf_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    bool owned = true;               // created by mrbf_comm_init (destroyed with the handle) vs. wrapped (mrbf_comm_from_nccl)
    cudaStream_t stream = nullptr;   // private stream for the collective (ordered against the caller by events)
    void* buf = nullptr; size_t cap = 0;      // staging: [counts (world ints, padded) | send rows | recv rows]
    char err[256] = {0};
};

namespace {
int cfail(mrbf_comm* c, int code, const char* what, const char* detail) {
    if (c) snprintf(c->err, sizeof(c->err), "%s%s%s", what, detail ? ": " : "", detail ? detail : "");
    return code;
}
int grow(mrbf_comm* c, size_t bytes) {
    if (c->cap >= bytes) return MRBF_OK;
    if (c->buf) { cudaStreamSynchronize(c->stream); cudaFree(c->buf); c->buf = nullptr; c->cap = 0; }
    cudaError_t e = cudaMalloc(&c->buf, bytes + bytes / 4);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return cfail(c, MRBF_ENOMEM, "cudaMalloc failed", cudaGetErrorString(e)); }
    c->cap = bytes + bytes / 4;
    return MRBF_OK;
}
}  // namespace

extern "C" {

int mrbf_comm_unique_id(char* id128) {
    if (!id128) return MRBF_EINVAL;
    NcclApi* a = nccl_api();
    if (!a) return MRBF_EUNSUPPORTED;
    ncclUniqueId id;
    if (a->GetUniqueId(&id) != 0) return MRBF_ECUDA;
    memcpy(id128, id.internal, 128);
    return MRBF_OK;
}

int mrbf_comm_init(int device, const char* id128, int32_t rank, int32_t world, mrbf_comm** out) {
    if (!out) return MRBF_EINVAL;
    *out = nullptr;
    if (!id128 || world < 1 || rank < 0 || rank >= world) return MRBF_EINVAL;
    NcclApi* a = nccl_api();
    if (!a) return MRBF_EUNSUPPORTED;
    if (cudaSetDevice(device) != cudaSuccess) return MRBF_ECUDA;
    mrbf_comm* c = new (std::nothrow) mrbf_comm();
    if (!c) return MRBF_ENOMEM;
    c->rank = rank; c->world = world; c->device = device;
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    if (a->CommInitRank(&c->comm, world, id, rank) != 0 || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        if (c->comm) a->CommDestroy(c->comm);
        delete c;
        return MRBF_ECUDA;
    }
    *out = c;
    return MRBF_OK;
}

int mrbf_comm_from_nccl(int device, void* nccl_comm, int32_t rank, int32_t world, mrbf_comm** out) {
    if (!out) return MRBF_EINVAL;
    *out = nullptr;
    if (!nccl_comm || world < 1 || rank < 0 || rank >= world) return MRBF_EINVAL;
    if (!nccl_api()) return MRBF_EUNSUPPORTED;
    if (cudaSetDevice(device) != cudaSuccess) return MRBF_ECUDA;
    mrbf_comm* c = new (std::nothrow) mrbf_comm();
    if (!c) return MRBF_ENOMEM;
    c->comm = (ncclComm_t)nccl_comm; c->rank = rank; c->world = world; c->device = device; c->owned = false;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return MRBF_ECUDA; }
    *out = c;
    return MRBF_OK;
}

void mrbf_comm_destroy(mrbf_comm* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    NcclApi* a = nccl_api();
    if (c->owned && c->comm && a) a->CommDestroy(c->comm);
    if (c->buf) cudaFree(c->buf);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* mrbf_comm_last_error(const mrbf_comm* c) {
    if (c) return c->err;
    NcclApi* a = nccl_api();
    return a ? "" : "NCCL could not be loaded (libnccl.so.2; set MRBF_NCCL_LIB)";
}

int mrbf_gather(mrbf_comm* c, const double* rows, int32_t count, int32_t width, int32_t max_count, double* all_rows, int32_t* counts) {
    if (!c || !all_rows || !counts || count < 0 || width <= 0 || max_count < count || (count > 0 && !rows)) return MRBF_EINVAL;
    NcclApi* a = nccl_api();
    if (!a) return MRBF_EUNSUPPORTED;
    if (cudaSetDevice(c->device) != cudaSuccess) return cfail(c, MRBF_ECUDA, "cudaSetDevice failed", nullptr);
    const int W = c->world;
    const size_t cnt_bytes = ((sizeof(int) * (size_t)(W + 1) + 255) / 256) * 256;
    const size_t row_bytes = sizeof(double) * (size_t)max_count * width;
    int rc = grow(c, cnt_bytes + row_bytes * (size_t)(W + 1));
    if (rc != MRBF_OK) return rc;
    int* d_cnt_all = (int*)c->buf; int* d_cnt = d_cnt_all + W;
    double* d_send = (double*)((char*)c->buf + cnt_bytes); double* d_recv = d_send + (size_t)max_count * width;
    cudaError_t e = cudaMemcpyAsync(d_cnt, &count, sizeof(int), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_send, 0, row_bytes, c->stream);
    if (e == cudaSuccess && count > 0) e = cudaMemcpyAsync(d_send, rows, sizeof(double) * (size_t)count * width, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) return cfail(c, MRBF_ECUDA, "staging copy failed", cudaGetErrorString(e));
    ncclResult_t r = a->AllGather(d_cnt, d_cnt_all, 1, ncclInt32, c->comm, c->stream);
    if (r == 0) r = a->AllGather(d_send, d_recv, (size_t)max_count * width, ncclFloat64, c->comm, c->stream);
    if (r != 0) return cfail(c, MRBF_ECUDA, "ncclAllGather failed", a->GetErrorString ? a->GetErrorString(r) : nullptr);
    e = cudaMemcpyAsync(counts, d_cnt_all, sizeof(int) * (size_t)W, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(all_rows, d_recv, row_bytes * (size_t)W, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return cfail(c, MRBF_ECUDA, "result copy failed", cudaGetErrorString(e));
    for (int r_ = 0; r_ < W; ++r_)
        if (counts[r_] < 0 || counts[r_] > max_count) return cfail(c, MRBF_EINVAL, "a rank sent more rows than max_count", nullptr);
    return MRBF_OK;
}

}  // extern "C"
