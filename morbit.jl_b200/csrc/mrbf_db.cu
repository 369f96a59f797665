// Device-resident database bookkeeping for the lock-step multistart driver.
//   db_append_kernel      new_result!(db, x, y)          /root/reference/src/Databases.jl:174-183, 202-205
//   model_scatter_kernel  swapping a rebuilt model in     /root/reference/src/SurrogateContainer.jl:376-382
// Both are pure copies (HBM-bound, a few KB per instance); one CTA per instance, coalesced rows.
#include "mrbf_common.cuh"
#include "mrbf_kernels.h"

namespace mrbf {

namespace {

__global__ void __launch_bounds__(128) db_append_kernel(DbAppendParams P) {
    const int b = blockIdx.x;
    const int na = P.n_add[b];
    const int base = P.n_db[b];                    // every thread reads the old count before thread 0 replaces it
    __syncthreads();
    if (na <= 0) {
        if (threadIdx.x == 0) { P.first_id[b] = base + 1; if (P.status) P.status[b] = na < 0 ? 1 : 0; }
        return;
    }
    if (na > P.add_stride || base < 0 || base + na > P.db_stride) {      // capacity exceeded: nothing is written
        if (threadIdx.x == 0) { P.first_id[b] = 0; if (P.status) P.status[b] = 1; }
        return;
    }
    const double* src = P.new_sites + (size_t)b * P.add_stride * P.n;
    double* dst = P.sites + ((size_t)b * P.db_stride + base) * P.n;
    for (int i = threadIdx.x; i < na * P.n; i += blockDim.x) dst[i] = src[i];
    double* vdst = P.values + ((size_t)b * P.db_stride + base) * P.k;
    if (P.new_values) {
        const double* vsrc = P.new_values + (size_t)b * P.add_stride * P.k;
        for (int i = threadIdx.x; i < na * P.k; i += blockDim.x) vdst[i] = vsrc[i];
    } else {
        const double qnan = __longlong_as_double(0x7ff8000000000000LL);  // value-less result (unevaluated), Databases.jl:202-205
        for (int i = threadIdx.x; i < na * P.k; i += blockDim.x) vdst[i] = qnan;
    }
    if (threadIdx.x == 0) { P.first_id[b] = base + 1; P.n_db[b] = base + na; if (P.status) P.status[b] = 0; }
}

__global__ void __launch_bounds__(256) model_scatter_kernel(ModelScatterParams P) {
    const int s = blockIdx.x;
    const int t = P.map[s];
    if (t < 0 || t >= P.B_dst) return;
    // the destination batch may have room for more training points per instance than the source (dst_stride >= train_stride):
    // rows are copied one to one, the destination's surplus rows are cleared (w = 0 => they contribute nothing)
    const size_t nc = (size_t)P.train_stride * P.n, nw = (size_t)P.train_stride * P.k, nl = (size_t)P.pl * P.k;
    const size_t ncd = (size_t)P.dst_stride * P.n, nwd = (size_t)P.dst_stride * P.k;
    const double* c = P.src_centers + s * nc; double* cd = P.dst_centers + t * ncd;
    for (size_t i = threadIdx.x; i < ncd; i += blockDim.x) cd[i] = i < nc ? c[i] : 0.0;
    const double* w = P.src_w + s * nw; double* wd = P.dst_w + t * nwd;
    for (size_t i = threadIdx.x; i < nwd; i += blockDim.x) wd[i] = i < nw ? w[i] : 0.0;
    const double* l = P.src_lam + s * nl; double* ld = P.dst_lam + t * nl;
    for (size_t i = threadIdx.x; i < nl; i += blockDim.x) ld[i] = l[i];
    if (threadIdx.x == 0) { P.dst_N[t] = P.src_N[s]; P.dst_alpha2[t] = P.src_alpha2[s]; }
}

}  // namespace

cudaError_t launch_db_append(const DbAppendParams& P, cudaStream_t s) {
    db_append_kernel<<<P.B, 128, 0, s>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_model_scatter(const ModelScatterParams& P, cudaStream_t s) {
    model_scatter_kernel<<<P.S, 256, 0, s>>>(P);
    return cudaGetLastError();
}

}  // namespace mrbf
