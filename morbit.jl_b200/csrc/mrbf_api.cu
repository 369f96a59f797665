// C ABI (include/morbit_rbf.h): context, workspaces, host<->device staging, kernel dispatch.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <new>
#include <utility>

#include <nvtx3/nvToolsExt.h>

#include "mrbf_common.cuh"
#include "mrbf_kernels.h"

using namespace mrbf;

namespace mrbf {
// MRBF_POISON=<byte> (debugging aid; see dev_malloc below for the global-memory half): before every kernel of the library the shared
// memory of every SM is filled with that byte -- a kernel whose result depends on shared memory it has not written (typically a padded
// operand multiplied by zero: 0 x NaN) then fails its test instead of depending on what ran on that SM before.  Serialises the device.
static const int g_poison = getenv("MRBF_POISON") ? atoi(getenv("MRBF_POISON")) : -1;
__global__ void poison_smem_kernel(unsigned long long pattern, int words) {
    extern __shared__ unsigned long long poison_sm[];
    for (int i = threadIdx.x; i < words; i += blockDim.x) poison_sm[i] = pattern;
}
static void poison_shared_memory() {
    if (g_poison < 0) return;
    int dev = 0, sms = 0, optin = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    unsigned long long pat = (unsigned long long)(g_poison & 255); pat |= pat << 8; pat |= pat << 16; pat |= pat << 32;
    cudaDeviceSynchronize();
    cudaFuncSetAttribute(poison_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
    poison_smem_kernel<<<4 * sms, 256, optin>>>(pat, optin / 8);      // a CTA takes a whole SM: every SM gets (at least) one
    cudaDeviceSynchronize();
}
cudaError_t raise_dyn_smem_impl(const void* kernel, size_t bytes) {
    poison_shared_memory();
    if (bytes <= 48 * 1024) return cudaSuccess;
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> limit;      // (function, device) -> what the attribute was last raised to
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> guard(mu);
    size_t& cur = limit[std::make_pair(kernel, dev)];
    if (bytes <= cur) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}
}  // namespace mrbf

namespace {

constexpr size_t SMEM_LIMIT = 225 * 1024;   // of the 227 KB a CTA may opt in to on sm_100a

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct mrbf_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int64_t launches = 0;
    char err[512] = {0};
    bool prof = false;
    double isapprox_rtol = 1.4901161193847656e-08;   // sqrt(eps(Float64)); see mrbf_set_isapprox_rtol
    cudaEvent_t ev0[8] = {nullptr}, ev1[8] = {nullptr};
    bool ev_used[8] = {false};
    DevBuf ws[24];      // kernel workspaces (grow-only)
    DevBuf hb[32];      // staging for the host-pointer entry points
};

struct mrbf_prepared {
    int B, n, NM, p, db_stride, found_stride, r4_stride, cfg_degree, kernel;
    double shape;
    size_t fs_stride;
    double* fs = nullptr;
    int* ints = nullptr;        // elig[B], n_found[B], n_extra[B], n_r4[B], found[B*found_stride], r4[B*r4_stride]
    int *elig, *n_found, *n_extra, *n_r4, *found, *r4;
    int kind = 0;               // 0: round4_block_kernel layout (round4_fast_state_layout), 1: round4_elim_kernel layout,
                                // 2: no factorisation at all (optimized_sampling = false): every instance takes the general route
    SchurGeom geom{};
};

struct mrbf_model {
    int B, n, k, train_stride, p, deg;
    int kernel, ibeta;
    double sgn;
    int* N = nullptr;
    double* centers = nullptr;
    double* w = nullptr;
    double* lam = nullptr;
    double* alpha2 = nullptr;
    double* pack = nullptr;     // tiled centred copy for the DMMA sweep (n <= 64)
    int pack_s = 0, pack_nt = 0;
    size_t pack_tile_doubles = 0;
    bool pack_valid = false;    // built lazily by the first values-only evaluation
};

extern "C" void mrbf_free_prepared(mrbf_ctx* ctx, mrbf_prepared* kp);

namespace {

int fail(mrbf_ctx* c, int code, const char* fmt, const char* detail = "") {
    if (c) snprintf(c->err, sizeof(c->err), fmt, detail);
    return code;
}

#define CK(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) return fail(ctx, MRBF_ECUDA, "CUDA error: %s", cudaGetErrorString(e_)); \
    } while (0)

// event bracket around a kernel class (instrumentation only)
// MRBF_NVTX=1 additionally opens an NVTX range per kernel class (header-only nvtx3; a no-op without an attached tool), so that
// `ncu --nvtx --nvtx-include "mrbf:round4/"` or a timeline tool can address the phases by name.
const bool g_nvtx = getenv("MRBF_NVTX") && atoi(getenv("MRBF_NVTX")) != 0;
const char* const k_phase_names[8] = {"mrbf:rounds123", "mrbf:round4", "mrbf:gather", "mrbf:build", "mrbf:eval", "mrbf:round4_literal",
                                      "mrbf:build_prepared", "mrbf:round4_prefix"};
struct Timed {
    mrbf_ctx* c; int id;
    Timed(mrbf_ctx* ctx, int i) : c(ctx), id(i) {
        mrbf::poison_shared_memory();         // (MRBF_POISON only; kernels with static shared memory do not pass raise_dyn_smem)
        if (g_nvtx) nvtxRangePushA(k_phase_names[id & 7]);
        if (c->prof) cudaEventRecord(c->ev0[id], c->stream);
    }
    ~Timed() {
        if (c->prof) { cudaEventRecord(c->ev1[id], c->stream); c->ev_used[id] = true; }
        if (g_nvtx) nvtxRangePop();
    }
};

// Every device allocation of the library goes through here.  MRBF_POISON=<byte> (debugging aid, stands in for compute-sanitizer's
// initcheck) fills fresh allocations with that byte -- 255 gives NaN doubles / -1 ints, 127 gives 1.4e306 / 2139062143 -- so that a
// kernel consuming workspace it has not written shows up as a wrong result instead of depending on what the allocator handed back.
static cudaError_t dev_malloc(void** p, size_t bytes) {
    static const int poison = getenv("MRBF_POISON") ? atoi(getenv("MRBF_POISON")) : -1;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaSuccess && poison >= 0) { e = cudaMemset(*p, poison & 255, bytes); if (e == cudaSuccess) e = cudaDeviceSynchronize(); }
    return e;
}
template <class T> static cudaError_t dev_malloc(T** p, size_t bytes) { return dev_malloc((void**)p, bytes); }

int ensure(mrbf_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return MRBF_OK;
    if (b.p) { cudaStreamSynchronize(ctx->stream); cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8;
    cudaError_t e = dev_malloc(&b.p, want);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();        // the failure is reported here; do not leave it for the next launch check to trip over
        b.p = nullptr; return fail(ctx, MRBF_ENOMEM, "cudaMalloc failed: %s", cudaGetErrorString(e));
    }
    b.cap = want;
    return MRBF_OK;
}
#define ENSURE(buf, bytes) do { int r_ = ensure(ctx, (buf), (bytes)); if (r_ != MRBF_OK) return r_; } while (0)

// Resolve RbfConfig's kernel parameters (_get_kernel_params, RbfModel.jl:665-690; package defaults for NaN).
int resolve_radfn(mrbf_ctx* ctx, const mrbf_cfg* cfg, double shape, RadFn* rf, double* alpha, int* cpd) {
    const bool nan = !(shape == shape);
    rf->kernel = cfg->kernel; rf->ibeta = 0; rf->sgn = 1.0;
    double a = 1.0;
    switch (cfg->kernel) {
    case MRBF_GAUSSIAN: a = nan ? 1.0 : shape; *cpd = 0; break;
    case MRBF_INV_MULTIQUADRIC: a = nan ? 1.0 : shape; *cpd = 0; break;
    case MRBF_MULTIQUADRIC: a = nan ? 1.0 : shape; rf->sgn = -1.0; *cpd = 1; break;        // (-1)^ceil(1/2)
    case MRBF_CUBIC: {
        int beta = nan ? 3 : (int)shape;
        if (beta < 1 || beta % 2 == 0) return fail(ctx, MRBF_EINVAL, "cubic exponent must be a positive odd integer%s");
        rf->ibeta = beta; *cpd = (beta + 1) / 2;                                            // ceil(beta/2)
        rf->sgn = ((*cpd) & 1) ? -1.0 : 1.0;
        break;
    }
    case MRBF_THIN_PLATE_SPLINE: {
        int kk = nan ? 2 : (int)shape;
        if (kk < 1) return fail(ctx, MRBF_EINVAL, "thin plate spline order must be >= 1%s");
        rf->ibeta = kk; *cpd = kk + 1; rf->sgn = ((kk + 1) & 1) ? -1.0 : 1.0;
        break;
    }
    default: return fail(ctx, MRBF_EINVAL, "unknown kernel id%s");
    }
    if (!(a > 0.0)) return fail(ctx, MRBF_EINVAL, "shape parameter must be strictly positive%s");
    rf->alpha2 = a * a;
    *alpha = a;
    return MRBF_OK;
}

int check_cfg(mrbf_ctx* ctx, const mrbf_cfg* cfg) {
    if (!cfg) return fail(ctx, MRBF_EINVAL, "cfg is NULL%s");
    if (cfg->polynomial_degree < -1 || cfg->polynomial_degree > 1)
        return fail(ctx, MRBF_EUNSUPPORTED, "polynomial_degree must be -1, 0 or 1%s");
    if (cfg->use_max_points) return fail(ctx, MRBF_EUNSUPPORTED, "use_max_points draws random points on the host; not supported%s");
    if (!(cfg->theta_enlarge_1 >= 1.0 && cfg->theta_enlarge_2 >= 1.0)) return fail(ctx, MRBF_EINVAL, "theta_enlarge must be >= 1%s");
    if (!(cfg->theta_enlarge_1 * cfg->theta_pivot <= 1.0)) return fail(ctx, MRBF_EINVAL, "theta_pivot must be <= 1/theta_enlarge_1%s");
    return MRBF_OK;
}

int max_points_of(const mrbf_cfg* cfg, int n) {
    return cfg->max_model_points <= 0 ? ((n + 1) * (n + 2)) / 2 : cfg->max_model_points;
}

// Allocate (or reuse, when every shape matches) the caller's kept-factorisation handle.
int ensure_prepared(mrbf_ctx* ctx, mrbf_prepared** keep_out, const mrbf_cfg* cfg, int B, int n, int NM, int p, int db_stride,
                    int found_stride, int r4_stride, size_t fsd, int kind, const SchurGeom& geom) {
    if (*keep_out && (*keep_out)->B == B && (*keep_out)->n == n && (*keep_out)->NM == NM && (*keep_out)->p == p &&
        (*keep_out)->found_stride == found_stride && (*keep_out)->r4_stride == r4_stride && (*keep_out)->fs_stride == fsd &&
        (*keep_out)->kind == kind) {
        mrbf_prepared* kp = *keep_out;       // reuse the caller's handle (same shapes): no allocation on the hot path
        kp->db_stride = db_stride; kp->cfg_degree = cfg->polynomial_degree; kp->kernel = cfg->kernel; kp->shape = cfg->shape_parameter;
        kp->geom = geom;
        return MRBF_OK;
    }
    if (*keep_out) { mrbf_free_prepared(ctx, *keep_out); *keep_out = nullptr; }
    mrbf_prepared* kp = new (std::nothrow) mrbf_prepared();
    if (!kp) return fail(ctx, MRBF_ENOMEM, "out of host memory%s");
    kp->B = B; kp->n = n; kp->NM = NM; kp->p = p; kp->db_stride = db_stride; kp->found_stride = found_stride;
    kp->r4_stride = r4_stride; kp->cfg_degree = cfg->polynomial_degree; kp->kernel = cfg->kernel; kp->shape = cfg->shape_parameter;
    kp->fs_stride = fsd; kp->kind = kind; kp->geom = geom;
    const size_t ni = (size_t)B * 4 + (size_t)B * found_stride + (size_t)B * r4_stride;
    cudaError_t e1 = dev_malloc(&kp->fs, (size_t)B * (fsd ? fsd : 1) * sizeof(double));
    cudaError_t e2 = (e1 == cudaSuccess) ? dev_malloc(&kp->ints, ni * sizeof(int)) : e1;
    if (e2 != cudaSuccess) {
        (void)cudaGetLastError();
        cudaFree(kp->fs); delete kp;
        return fail(ctx, MRBF_ENOMEM, "cudaMalloc failed: %s", cudaGetErrorString(e2));
    }
    kp->elig = kp->ints; kp->n_found = kp->elig + B; kp->n_extra = kp->n_found + B; kp->n_r4 = kp->n_extra + B;
    kp->found = kp->n_r4 + B; kp->r4 = kp->found + (size_t)B * found_stride;
    *keep_out = kp;
    return MRBF_OK;
}

int run_round4(mrbf_ctx* ctx, const mrbf_cfg* cfg, int B, int n, int db_stride, const double* sites, const int* n_db,
               const double* lb2, const double* ub2, int found_stride, const int* found, const int* n_found,
               int extra_stride, const double* extra, const int* n_extra, int n0max, int r4_stride, int* r4, int* n_r4, int* status,
               mrbf_prepared** keep_out = nullptr, const unsigned char* cflags = nullptr) {
    RadFn rf; double alpha; int cpd;
    int rc = resolve_radfn(ctx, cfg, cfg->shape_parameter, &rf, &alpha, &cpd);
    if (rc != MRBF_OK) return rc;
    const int p = poly_dim(n, cfg->polynomial_degree);
    const int max_points = max_points_of(cfg, n);
    // N never exceeds max_points (loop bound) nor n0max + #candidates
    int NM = n0max + db_stride; if (NM > max_points) NM = max_points; if (NM < n0max) NM = n0max;
    Round4Params R{};
    R.B = B; R.n = n; R.db_stride = db_stride; R.found_stride = found_stride; R.extra_stride = extra_stride; R.r4_stride = r4_stride;
    R.NM = NM; R.max_points = max_points;
    R.cfg.polynomial_degree = cfg->polynomial_degree; R.cfg.optimized_sampling = cfg->optimized_sampling;
    R.rf = rf;
    const double t2 = cfg->theta_pivot_cholesky * cfg->theta_pivot_cholesky;
    R.chol_thr = t2 * t2;
    R.sites = sites; R.n_db = n_db; R.lb2 = lb2; R.ub2 = ub2; R.found = found; R.n_found = n_found;
    R.extra_sites = extra; R.n_extra = n_extra; R.r4 = r4; R.n_r4 = n_r4; R.status = status;
    ENSURE(ctx->ws[6], (size_t)B * db_stride * 5);      // int candidate list + byte flags per database entry
    R.cand = (unsigned char*)ctx->ws[6].p;
    R.cflags = cflags;
    // 1. fast paths for the regular case N0 == p; both mark the instances they cannot take with n_r4 = -1.
    //    (a) register-tiled right-looking elimination when the database has <= 128 sites and its panels fit in shared memory,
    //    (b) else the blocked left-looking kernel (state in shared memory or in a global workspace).
    {
        const SchurGeom geom = round4_schur_geom(n, p, db_stride);
        const bool schur = geom.eligible != 0;
        const size_t fsd = schur ? geom.state_doubles : round4_fast_state_doubles(n, NM, p);
        int Tb = 8;
        size_t fv = 0, fsmem = 0;
        if (!schur) {
            // block size: 8 candidates per block when the block buffers still fit beside the state in shared memory, else 4
            fv = round4_block_vec_doubles(8, n, NM, p);
            if ((fv + fsd) * sizeof(double) > SMEM_LIMIT && (round4_block_vec_doubles(4, n, NM, p) + fsd) * sizeof(double) <= SMEM_LIMIT) {
                Tb = 4; fv = round4_block_vec_doubles(4, n, NM, p);
            }
            fsmem = fv * sizeof(double);
            if (fsmem > SMEM_LIMIT) return fail(ctx, MRBF_EUNSUPPORTED, "max_model_points too large for the round-4 kernel%s");
        }
        R.fs_stride = fsd;
        if (keep_out) {
            rc = ensure_prepared(ctx, keep_out, cfg, B, n, NM, p, db_stride, found_stride, r4_stride, fsd, schur ? 1 : 0, geom);
            if (rc != MRBF_OK) return rc;
            R.keep_fs = (*keep_out)->fs; R.elig = (*keep_out)->elig;
        }
        if (schur) {
            // Under-poised instances (N0 < p) first: the literal kernel walks them until the point set is poised and hands them over to
            // the register kernels below (Round4Params::hyb).  Regular instances leave that launch at once.
            // Only for kernels that are conditionally positive definite of order <= 1 (Gaussian, inverse multiquadric, multiquadric,
            // cubic with exponent 1).  The reference appends a column to Z on EVERY acceptance (RbfModel.jl:464-467), also while N < p,
            // so at N = p its Z holds p - N0 directions that are orthogonal to the constants (the first column of Q spans them) but not to
            // the linear polynomials.  For order <= 1 the reduced kernel matrix is positive definite on any such Z: tau^2 > 0 for every
            // candidate that is not a duplicate, exactly like the walk that starts afresh from the poised set -- same decisions.  For
            // order 2 (cubic, thin plate spline) it is indefinite there and the reference goes on rejecting candidates a fresh walk
            // accepts (tests/test_handover_property.py holds the counterexample): those stay on the literal kernel for the whole walk.
            // The prefix run never holds more than p points: with NM = p its whole state (Phi, Q, R, Z, L^-1) lives in shared memory.
            const size_t lvec = round4_vec_doubles(n, p, p), lws = round4_ws_doubles(n, p, p);
            if (p > 0 && cpd <= 1 && (lvec + lws) * sizeof(double) <= SMEM_LIMIT) {
                ENSURE(ctx->ws[19], sizeof(int) * 3 * (size_t)B);
                R.hyb = (int*)ctx->ws[19].p; R.pre_cnt = R.hyb + B; R.pre_min = R.pre_cnt + B;
                Round4Params Rp = R;
                Rp.NM = p;
                Rp.prefix_mode = 1; Rp.only_marked = 0; Rp.ws_in_smem = 1; Rp.ws = nullptr; Rp.ws_stride = 0; Rp.b0 = 0;
                Timed t_(ctx, 7);
                CK(launch_round4(Rp, (lvec + lws) * sizeof(double), ctx->stream, B));
                ctx->launches += 1;
            }
            ENSURE(ctx->ws[13], (size_t)B * geom.pw_doubles * sizeof(double));
            R.panel_ws = (double*)ctx->ws[13].p;
            const bool dbg = getenv("MRBF_DEBUG_CLOCK") != nullptr;
            if (dbg) { ENSURE(ctx->ws[12], 1024 * sizeof(long long)); CK(cudaMemsetAsync(ctx->ws[12].p, 0, 1024 * sizeof(long long), ctx->stream)); R.dbg_clock = (long long*)ctx->ws[12].p; }
            {
                Timed t_(ctx, 1);
                CK(launch_round4_schur(R, geom, ctx->stream));
            }
            if (dbg) {
                long long h[1024];
                CK(cudaMemcpyAsync(h, ctx->ws[12].p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream));
                fprintf(stderr, "[mrbf clock] prep: setup %lld GJ %lld | elim: load %lld panels %lld tiles %lld elim %lld | blocks (start->panel published):", h[1] - h[0],
                        h[2] - h[1], h[100] - h[6], h[7] - h[100], h[4] - h[7], h[5] - h[4]);
                fprintf(stderr, " [tiles: distances %lld phi %lld C.V %lld]", h[101] - h[7], h[102] - h[101], h[4] - h[102]);
                for (int K = 0; K < 40 && h[8 + 2 * K]; ++K) fprintf(stderr, " %lld/%lld", h[8 + 2 * K] - h[4], h[9 + 2 * K] ? h[9 + 2 * K] - h[8 + 2 * K] : -1LL);
                fprintf(stderr, "\n[mrbf clock] per block: diag pivots / publish+arrive / panel wakes after arrive / panel work / next diag tile: panel published / updated / starts:");
                for (int K = 0; K < 25 && h[128 + 8 * K]; ++K)
                    fprintf(stderr, " %lld/%lld/%lld/%lld/%lld/%lld/%lld", h[129 + 8 * K] - h[128 + 8 * K], h[130 + 8 * K] - h[129 + 8 * K], h[131 + 8 * K] - h[130 + 8 * K],
                            h[132 + 8 * K] - h[131 + 8 * K], h[133 + 8 * K] - h[132 + 8 * K], h[134 + 8 * K] - h[133 + 8 * K], h[136 + 8 * K] - h[134 + 8 * K]);
                fprintf(stderr, "\n[mrbf clock] arrival of every warp at the panel barrier, cycles after the diagonal tile's hand-over:");
                for (int K = 1; K < 25 && h[128 + 8 * K]; ++K) {
                    fprintf(stderr, " [K=%d", K);
                    for (int w = 0; w < 12; ++w) if (h[600 + 12 * K + w]) fprintf(stderr, " %lld", h[600 + 12 * K + w] - h[130 + 8 * K]);
                    fprintf(stderr, "]");
                }
                fprintf(stderr, "\n[mrbf clock] leverage warp per block: after diag(K-1) -> u vectors / pass over M / reduce + hand-over; lead over the diagonal tile's start:");
                for (int K = 1; K < 25 && h[340 + 4 * K]; ++K)
                    fprintf(stderr, " %lld/%lld/%lld;%lld", h[341 + 4 * K] - h[340 + 4 * K], h[342 + 4 * K] - h[341 + 4 * K], h[343 + 4 * K] - h[342 + 4 * K],
                            h[128 + 8 * K] - h[343 + 4 * K]);
                fprintf(stderr, "\n");
            }
        } else {
            if ((fv + fsd) * sizeof(double) <= SMEM_LIMIT) { R.fs_in_smem = 1; fsmem = (fv + fsd) * sizeof(double); R.fs = nullptr; }
            else {
                R.fs_in_smem = 0;
                // the state streams from L2 / HBM: when the reduced system can get large (many model points beyond the polynomial
                // basis) 16 candidates share every pass over the packed L^{-1}
                const char* t16 = getenv("MRBF_R4_T16");
                const bool want16 = t16 ? atoi(t16) != 0 : (NM - p >= 192);
                if (want16 && round4_block_vec_doubles(16, n, NM, p) * sizeof(double) <= SMEM_LIMIT) {
                    Tb = 16; fsmem = round4_block_vec_doubles(16, n, NM, p) * sizeof(double);
                }
                if (!keep_out) { ENSURE(ctx->ws[9], (size_t)B * fsd * sizeof(double)); R.fs = (double*)ctx->ws[9].p; }
            }
            Timed t_(ctx, 1);
            CK(launch_round4_block(R, Tb, fsmem, ctx->stream));
        }
        ctx->launches += 1;
    }
    // 2. literal kernel for the marked rest (N0 != p: budget-limited round 3, explicit found sets, rank-deficient Pi_0)
    const size_t vecd = round4_vec_doubles(n, NM, p), wsd = round4_ws_doubles(n, NM, p);
    size_t smem = vecd * sizeof(double);
    R.only_marked = 1;
    if ((vecd + wsd) * sizeof(double) <= SMEM_LIMIT) {
        R.ws_in_smem = 1; smem = (vecd + wsd) * sizeof(double); R.ws = nullptr; R.ws_stride = 0; R.b0 = 0;
        Timed t_(ctx, 5);
        CK(launch_round4(R, smem, ctx->stream, B));
        ctx->launches += 1;
    } else {
        if (smem > SMEM_LIMIT) return fail(ctx, MRBF_EUNSUPPORTED, "max_model_points too large for the round-4 kernel%s");
        R.ws_in_smem = 0; R.ws_stride = wsd;
        size_t chunk = ((size_t)4 << 30) / (wsd * sizeof(double));       // <= 4 GiB of workspace per launch
        if (chunk < 1) chunk = 1;
        if (chunk > (size_t)B) chunk = B;
        ENSURE(ctx->ws[7], chunk * wsd * sizeof(double));
        R.ws = (double*)ctx->ws[7].p;
        Timed t_(ctx, 5);
        for (int b0 = 0; b0 < B; b0 += (int)chunk) {
            R.b0 = b0;
            const int g = (B - b0) < (int)chunk ? (B - b0) : (int)chunk;
            CK(launch_round4(R, smem, ctx->stream, g));
            ctx->launches += 1;
        }
    }
    if (keep_out) {                           // ids needed later to gather the training values in training order
        mrbf_prepared* kp = *keep_out;
        const size_t sB = sizeof(int) * (size_t)B;
        CK(cudaMemcpyAsync(kp->n_found, n_found, sB, cudaMemcpyDeviceToDevice, ctx->stream));
        if (n_extra) CK(cudaMemcpyAsync(kp->n_extra, n_extra, sB, cudaMemcpyDeviceToDevice, ctx->stream));
        else CK(cudaMemsetAsync(kp->n_extra, 0, sB, ctx->stream));
        CK(cudaMemcpyAsync(kp->n_r4, n_r4, sB, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(kp->found, found, sB * found_stride, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaMemcpyAsync(kp->r4, r4, sB * r4_stride, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return MRBF_OK;
}

}  // namespace

extern "C" {

int mrbf_abi_version(void) { return MRBF_ABI_VERSION; }

int mrbf_init(int device, mrbf_ctx** out) {
    if (!out) return MRBF_EINVAL;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return MRBF_ECUDA;
    mrbf_ctx* ctx = new (std::nothrow) mrbf_ctx();
    if (!ctx) return MRBF_ENOMEM;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return MRBF_ECUDA;
    }
    ctx->own_stream = true;
    *out = ctx;
    return MRBF_OK;
}

int mrbf_set_stream(mrbf_ctx* ctx, void* s) {
    if (!ctx) return MRBF_EINVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) { cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
    ctx->stream = (cudaStream_t)s;
    return MRBF_OK;
}

int mrbf_get_stream(const mrbf_ctx* ctx, void** s) {
    if (!ctx || !s) return MRBF_EINVAL;
    *s = (void*)ctx->stream;
    return MRBF_OK;
}

int mrbf_set_isapprox_rtol(mrbf_ctx* ctx, double rtol) {
    if (!ctx || !(rtol >= 0.0)) return MRBF_EINVAL;
    ctx->isapprox_rtol = rtol;
    return MRBF_OK;
}

int mrbf_sync(mrbf_ctx* ctx) {
    if (!ctx) return MRBF_EINVAL;
    CK(cudaStreamSynchronize(ctx->stream));
    return MRBF_OK;
}

void mrbf_destroy(mrbf_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& b : ctx->ws) if (b.p) cudaFree(b.p);
    for (auto& b : ctx->hb) if (b.p) cudaFree(b.p);
    for (int i = 0; i < 8; ++i) { if (ctx->ev0[i]) cudaEventDestroy(ctx->ev0[i]); if (ctx->ev1[i]) cudaEventDestroy(ctx->ev1[i]); }
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int mrbf_profile_enable(mrbf_ctx* ctx, int32_t on) {
    if (!ctx) return MRBF_EINVAL;
    CK(cudaSetDevice(ctx->device));
    if (on && !ctx->ev0[0])
        for (int i = 0; i < 8; ++i) { CK(cudaEventCreate(&ctx->ev0[i])); CK(cudaEventCreate(&ctx->ev1[i])); }
    ctx->prof = on != 0;
    for (int i = 0; i < 8; ++i) ctx->ev_used[i] = false;
    return MRBF_OK;
}

int mrbf_profile_read(mrbf_ctx* ctx, double* ms8) {
    if (!ctx || !ms8) return MRBF_EINVAL;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 8; ++i) {
        ms8[i] = 0.0;
        if (ctx->ev_used[i]) { float ms = 0.f; CK(cudaEventElapsedTime(&ms, ctx->ev0[i], ctx->ev1[i])); ms8[i] = ms; }
    }
    return MRBF_OK;
}

const char* mrbf_last_error(const mrbf_ctx* ctx) { return ctx ? ctx->err : "null context"; }
int64_t mrbf_launch_count(const mrbf_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------------------------------------ select
static int select_points_impl(mrbf_ctx* ctx, const mrbf_cfg* cfg, int32_t B, int32_t n, int32_t db_stride,
                           const double* sites, const int32_t* n_db, const int32_t* x_index, const double* x,
                           const double* delta, double delta_max, const double* glb, const double* gub,
                           const int32_t* flags_in, const int32_t* max_new,
                           int32_t* r1, int32_t* n_r1, int32_t* r2, int32_t* n_r2, double* r3_sites, int32_t* n_r3,
                           int32_t r4_stride, int32_t* r4, int32_t* n_r4, double* dirs, int32_t* n_dirs,
                           int32_t* flags_out, int32_t* status, mrbf_prepared** keep) {
    if (!ctx) return MRBF_EINVAL;
    int rc = check_cfg(ctx, cfg);
    if (rc != MRBF_OK) return rc;
    if (B <= 0 || n <= 0 || db_stride <= 0) return fail(ctx, MRBF_EINVAL, "B, n and db_stride must be positive%s");
    CK(cudaSetDevice(ctx->device));
    SelectParams S{};
    S.B = B; S.n = n; S.db_stride = db_stride; S.found_stride = 2 * n + 1;
    S.cfg.polynomial_degree = cfg->polynomial_degree; S.cfg.optimized_sampling = cfg->optimized_sampling;
    S.cfg.theta_enlarge_1 = cfg->theta_enlarge_1; S.cfg.theta_enlarge_2 = cfg->theta_enlarge_2; S.cfg.theta_pivot = cfg->theta_pivot;
    S.delta_max = delta_max; S.approx_rtol = ctx->isapprox_rtol;
    S.sites = sites; S.n_db = n_db; S.x_index = x_index; S.x = x; S.delta = delta; S.glb = glb; S.gub = gub;
    S.flags_in = flags_in; S.max_new = max_new;
    S.r1 = r1; S.n_r1 = n_r1; S.r2 = r2; S.n_r2 = n_r2; S.r3_sites = r3_sites; S.n_r3 = n_r3; S.dirs = dirs; S.n_dirs = n_dirs;
    S.flags_out = flags_out;
    const int ldz = (n + 3) & ~3;
    S.wz_in_smem = select_smem_bytes(n, true, 0, db_stride) <= SMEM_LIMIT;
    // projection coefficients (n x ldS doubles, ldS = db_stride rounded up to even) stay in shared memory when three CTAs
    // still fit per SM; the shifted seeds always live in the global workspace
    const int ldS = (db_stride + 1) & ~1;
    const int st_doubles = n * ldS;
    S.st_in_smem = S.wz_in_smem && select_smem_bytes(n, true, st_doubles, db_stride) <= 74 * 1024;
    const size_t seeds = (size_t)B * ldS * n * sizeof(double);
    ENSURE(ctx->ws[0], seeds); ENSURE(ctx->ws[1], S.st_in_smem ? 16 : seeds); ENSURE(ctx->ws[2], (size_t)B * db_stride);
    S.S = (double*)ctx->ws[0].p; S.T = (double*)ctx->ws[1].p; S.cflags = (unsigned char*)ctx->ws[2].p;
    if (!S.wz_in_smem) { ENSURE(ctx->ws[3], (size_t)B * 2 * n * ldz * sizeof(double)); S.WZ = (double*)ctx->ws[3].p; }
    ENSURE(ctx->ws[4], (size_t)B * n * 2 * sizeof(double));
    S.lb2 = (double*)ctx->ws[4].p; S.ub2 = S.lb2 + (size_t)B * n;
    ENSURE(ctx->ws[5], ((size_t)B * S.found_stride + B) * sizeof(int));
    S.found = (int*)ctx->ws[5].p; S.n_found = S.found + (size_t)B * S.found_stride;
    const bool dbg123 = getenv("MRBF_DEBUG_CLOCK") != nullptr;
    if (dbg123) { ENSURE(ctx->ws[14], 64 * sizeof(long long)); CK(cudaMemsetAsync(ctx->ws[14].p, 0, 64 * sizeof(long long), ctx->stream)); S.dbg_clock = (long long*)ctx->ws[14].p; }
    // n <= 32 and <= 128 sites per database (the C3 shape): filter state in registers, scores on the FP64 tensor path
    static const bool mma_off = getenv("MRBF_SELECT_MMA") && atoi(getenv("MRBF_SELECT_MMA")) == 0;
    const bool use_mma = !mma_off && !dbg123 && select_mma_eligible(n, db_stride);
    {
        Timed t_(ctx, 0);
        if (use_mma) CK(launch_select_rounds123_mma(S, ctx->stream));
        else CK(launch_select_rounds123(S, select_smem_bytes(n, S.wz_in_smem, S.st_in_smem ? st_doubles : 0, db_stride), ctx->stream));
    }
    ctx->launches += 1;
    if (dbg123) {
        long long h[64];
        CK(cudaMemcpyAsync(h, ctx->ws[14].p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        fprintf(stderr, "[mrbf clock] rounds 1-3 of instance 0: %lld cycles; first filter steps (dlarfg / reflector update / W D^-2 / scores / arg-max):", h[1] - h[0]);
        for (int q = 0; q < 4 && h[8 + 8 * q]; ++q)
            fprintf(stderr, " %lld/%lld/%lld/%lld/%lld", h[9 + 8 * q] - h[8 + 8 * q], h[10 + 8 * q] - h[9 + 8 * q], h[11 + 8 * q] - h[10 + 8 * q],
                    h[12 + 8 * q] - h[11 + 8 * q], h[13 + 8 * q] - h[12 + 8 * q]);
        fprintf(stderr, "\n");
    }
    if (cfg->optimized_sampling) {           // RbfModel.jl:647-652
        rc = run_round4(ctx, cfg, B, n, db_stride, sites, n_db, S.lb2, S.ub2, S.found_stride, S.found, S.n_found,
                        n, r3_sites, n_r3, n + 1, r4_stride, r4, n_r4, status, keep, S.cflags);
        if (rc != MRBF_OK) return rc;
    } else {
        CK(cudaMemsetAsync(n_r4, 0, sizeof(int) * (size_t)B, ctx->stream));
        if (status) CK(cudaMemsetAsync(status, 0, sizeof(int) * (size_t)B, ctx->stream));
        if (keep) {
            // No round 4, hence no factorisation (RbfModel.jl:564-569, 647-652): the handle still records the found set so that
            // mrbf_build_prepared* builds [centre; r1; r2; r3] by the general route (elig = 0, n_r4 = 0 for every instance).
            const int p = poly_dim(n, cfg->polynomial_degree);
            int NM = n + 1 + db_stride; const int mpts = max_points_of(cfg, n); if (NM > mpts) NM = mpts; if (NM < n + 1) NM = n + 1;
            rc = ensure_prepared(ctx, keep, cfg, B, n, NM, p, db_stride, S.found_stride, r4_stride, 0, 2, SchurGeom{});
            if (rc != MRBF_OK) return rc;
            mrbf_prepared* kp = *keep;
            const size_t sB = sizeof(int) * (size_t)B;
            CK(cudaMemsetAsync(kp->elig, 0, sB, ctx->stream));
            CK(cudaMemcpyAsync(kp->n_found, S.n_found, sB, cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaMemsetAsync(kp->n_extra, 0, sB, ctx->stream));
            CK(cudaMemsetAsync(kp->n_r4, 0, sB, ctx->stream));
            CK(cudaMemcpyAsync(kp->found, S.found, sB * S.found_stride, cudaMemcpyDeviceToDevice, ctx->stream));
            if (r4_stride > 0) CK(cudaMemsetAsync(kp->r4, 0, sB * r4_stride, ctx->stream));
        }
    }
    return MRBF_OK;
}

int mrbf_select_points_dev(mrbf_ctx* ctx, const mrbf_cfg* cfg, int32_t B, int32_t n, int32_t db_stride,
                           const double* sites, const int32_t* n_db, const int32_t* x_index, const double* x,
                           const double* delta, double delta_max, const double* glb, const double* gub,
                           const int32_t* flags_in, const int32_t* max_new,
                           int32_t* r1, int32_t* n_r1, int32_t* r2, int32_t* n_r2, double* r3_sites, int32_t* n_r3,
                           int32_t r4_stride, int32_t* r4, int32_t* n_r4, double* dirs, int32_t* n_dirs,
                           int32_t* flags_out, int32_t* status) {
    return select_points_impl(ctx, cfg, B, n, db_stride, sites, n_db, x_index, x, delta, delta_max, glb, gub, flags_in, max_new,
                              r1, n_r1, r2, n_r2, r3_sites, n_r3, r4_stride, r4, n_r4, dirs, n_dirs, flags_out, status, nullptr);
}

int mrbf_select_points_keep_dev(mrbf_ctx* ctx, const mrbf_cfg* cfg, int32_t B, int32_t n, int32_t db_stride,
                                const double* sites, const int32_t* n_db, const int32_t* x_index, const double* x,
                                const double* delta, double delta_max, const double* glb, const double* gub,
                                const int32_t* flags_in, const int32_t* max_new,
                                int32_t* r1, int32_t* n_r1, int32_t* r2, int32_t* n_r2, double* r3_sites, int32_t* n_r3,
                                int32_t r4_stride, int32_t* r4, int32_t* n_r4, double* dirs, int32_t* n_dirs,
                                int32_t* flags_out, int32_t* status, mrbf_prepared** prepared) {
    if (!prepared) return MRBF_EINVAL;
    return select_points_impl(ctx, cfg, B, n, db_stride, sites, n_db, x_index, x, delta, delta_max, glb, gub, flags_in, max_new,
                              r1, n_r1, r2, n_r2, r3_sites, n_r3, r4_stride, r4, n_r4, dirs, n_dirs, flags_out, status, prepared);
}

void mrbf_free_prepared(mrbf_ctx* ctx, mrbf_prepared* kp) {
    if (!kp) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    cudaFree(kp->fs); cudaFree(kp->ints);
    delete kp;
}

#define H2D(dst, src, bytes) CK(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyHostToDevice, ctx->stream))
#define D2H(dst, src, bytes) CK(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDeviceToHost, ctx->stream))

static int select_points_host(mrbf_ctx* ctx, const mrbf_cfg* cfg, int32_t B, int32_t n, int32_t db_stride,
                       const double* sites, const int32_t* n_db, const int32_t* x_index, const double* x,
                       const double* delta, double delta_max, const double* glb, const double* gub,
                       const int32_t* flags_in, const int32_t* max_new,
                       int32_t* r1, int32_t* n_r1, int32_t* r2, int32_t* n_r2, double* r3_sites, int32_t* n_r3,
                       int32_t r4_stride, int32_t* r4, int32_t* n_r4, double* dirs, int32_t* n_dirs,
                       int32_t* flags_out, int32_t* status, mrbf_prepared** keep) {
    if (!ctx) return MRBF_EINVAL;
    if (B <= 0 || n <= 0 || db_stride <= 0 || r4_stride < 0) return fail(ctx, MRBF_EINVAL, "bad sizes%s");
    CK(cudaSetDevice(ctx->device));
    const size_t sB = sizeof(int) * (size_t)B, dBn = sizeof(double) * (size_t)B * n;
    const size_t sites_b = sizeof(double) * (size_t)B * db_stride * n;
    ENSURE(ctx->hb[0], sites_b); ENSURE(ctx->hb[1], sB * 5 + sB * 2); ENSURE(ctx->hb[2], dBn + sizeof(double) * (size_t)B + 2 * sizeof(double) * n);
    ENSURE(ctx->hb[3], sizeof(int) * ((size_t)B * n * 2 + (size_t)B * (size_t)(r4_stride > 0 ? r4_stride : 1) + (size_t)B * 8));
    ENSURE(ctx->hb[4], sizeof(double) * (size_t)B * n * n * 2);
    double* d_sites = (double*)ctx->hb[0].p;
    int* d_ndb = (int*)ctx->hb[1].p; int* d_xi = d_ndb + B; int* d_maxnew = d_xi + B; int* d_flags = d_maxnew + B;   // flags: 2B
    double* d_x = (double*)ctx->hb[2].p; double* d_delta = d_x + (size_t)B * n; double* d_glb = d_delta + B; double* d_gub = d_glb + n;
    int* d_r1 = (int*)ctx->hb[3].p; int* d_r2 = d_r1 + (size_t)B * n; int* d_r4 = d_r2 + (size_t)B * n;
    int* d_cnt = d_r4 + (size_t)B * (r4_stride > 0 ? r4_stride : 1);   // n_r1,n_r2,n_r3,n_r4,n_dirs,(flags_out 2),status
    double* d_r3 = (double*)ctx->hb[4].p; double* d_dirs = d_r3 + (size_t)B * n * n;
    H2D(d_sites, sites, sites_b); H2D(d_ndb, n_db, sB); H2D(d_xi, x_index, sB); H2D(d_maxnew, max_new, sB); H2D(d_flags, flags_in, 2 * sB);
    H2D(d_x, x, dBn); H2D(d_delta, delta, sizeof(double) * (size_t)B); H2D(d_glb, glb, sizeof(double) * n); H2D(d_gub, gub, sizeof(double) * n);
    int rc = select_points_impl(ctx, cfg, B, n, db_stride, d_sites, d_ndb, d_xi, d_x, d_delta, delta_max, d_glb, d_gub, d_flags,
                                d_maxnew, d_r1, d_cnt, d_r2, d_cnt + B, d_r3, d_cnt + 2 * B, r4_stride, d_r4, d_cnt + 3 * B,
                                d_dirs, d_cnt + 4 * B, d_cnt + 5 * B, d_cnt + 7 * B, keep);
    if (rc != MRBF_OK) return rc;
    D2H(r1, d_r1, sizeof(int) * (size_t)B * n); D2H(r2, d_r2, sizeof(int) * (size_t)B * n);
    if (r4_stride > 0) D2H(r4, d_r4, sizeof(int) * (size_t)B * r4_stride);
    D2H(n_r1, d_cnt, sB); D2H(n_r2, d_cnt + B, sB); D2H(n_r3, d_cnt + 2 * B, sB); D2H(n_r4, d_cnt + 3 * B, sB);
    D2H(n_dirs, d_cnt + 4 * B, sB); D2H(flags_out, d_cnt + 5 * B, 2 * sB);
    if (status) D2H(status, d_cnt + 7 * B, sB);
    D2H(r3_sites, d_r3, sizeof(double) * (size_t)B * n * n); D2H(dirs, d_dirs, sizeof(double) * (size_t)B * n * n);
    CK(cudaStreamSynchronize(ctx->stream));
    return MRBF_OK;
}

int mrbf_select_points(mrbf_ctx* ctx, const mrbf_cfg* cfg, int32_t B, int32_t n, int32_t db_stride,
                       const double* sites, const int32_t* n_db, const int32_t* x_index, const double* x,
                       const double* delta, double delta_max, const double* glb, const double* gub,
                       const int32_t* flags_in, const int32_t* max_new,
                       int32_t* r1, int32_t* n_r1, int32_t* r2, int32_t* n_r2, double* r3_sites, int32_t* n_r3,
                       int32_t r4_stride, int32_t* r4, int32_t* n_r4, double* dirs, int32_t* n_dirs,
                       int32_t* flags_out, int32_t* status) {
    return select_points_host(ctx, cfg, B, n, db_stride, sites, n_db, x_index, x, delta, delta_max, glb, gub, flags_in, max_new,
                              r1, n_r1, r2, n_r2, r3_sites, n_r3, r4_stride, r4, n_r4, dirs, n_dirs, flags_out, status, nullptr);
}

int mrbf_select_points_keep(mrbf_ctx* ctx, const mrbf_cfg* cfg, int32_t B, int32_t n, int32_t db_stride,
                            const double* sites, const int32_t* n_db, const int32_t* x_index, const double* x,
                            const double* delta, double delta_max, const double* glb, const double* gub,
                            const int32_t* flags_in, const int32_t* max_new,
                            int32_t* r1, int32_t* n_r1, int32_t* r2, int32_t* n_r2, double* r3_sites, int32_t* n_r3,
                            int32_t r4_stride, int32_t* r4, int32_t* n_r4, double* dirs, int32_t* n_dirs,
                            int32_t* flags_out, int32_t* status, mrbf_prepared** prepared) {
    if (!prepared) return MRBF_EINVAL;
    return select_points_host(ctx, cfg, B, n, db_stride, sites, n_db, x_index, x, delta, delta_max, glb, gub, flags_in, max_new,
                              r1, n_r1, r2, n_r2, r3_sites, n_r3, r4_stride, r4, n_r4, dirs, n_dirs, flags_out, status, prepared);
}

int mrbf_round4(mrbf_ctx* ctx, const mrbf_cfg* cfg, int32_t B, int32_t n, int32_t db_stride,
                const double* sites, const int32_t* n_db, const double* lb2, const double* ub2,
                int32_t found_stride, const int32_t* found, const int32_t* n_found,
                int32_t extra_stride, const double* extra_sites, const int32_t* n_extra,
                int32_t r4_stride, int32_t* r4, int32_t* n_r4, int32_t* status) {
    if (!ctx) return MRBF_EINVAL;
    int rc = check_cfg(ctx, cfg);
    if (rc != MRBF_OK) return rc;
    if (B <= 0 || n <= 0 || db_stride <= 0 || found_stride <= 0 || r4_stride <= 0) return fail(ctx, MRBF_EINVAL, "bad sizes%s");
    if (!extra_sites) extra_stride = 0;
    CK(cudaSetDevice(ctx->device));
    const size_t sB = sizeof(int) * (size_t)B;
    const size_t sites_b = sizeof(double) * (size_t)B * db_stride * n;
    const size_t extra_b = sizeof(double) * (size_t)B * extra_stride * n;
    ENSURE(ctx->hb[0], sites_b);
    ENSURE(ctx->hb[5], sizeof(double) * (size_t)B * n * 2 + extra_b);
    ENSURE(ctx->hb[6], sB * 4 + sizeof(int) * (size_t)B * ((size_t)found_stride + r4_stride));
    double* d_sites = (double*)ctx->hb[0].p;
    double* d_lb = (double*)ctx->hb[5].p; double* d_ub = d_lb + (size_t)B * n; double* d_extra = d_ub + (size_t)B * n;
    int* d_ndb = (int*)ctx->hb[6].p; int* d_nf = d_ndb + B; int* d_ne = d_nf + B; int* d_nr4 = d_ne + B;
    int* d_found = d_nr4 + B; int* d_r4 = d_found + (size_t)B * found_stride;
    ENSURE(ctx->hb[7], sB);
    int* d_status = (int*)ctx->hb[7].p;
    H2D(d_sites, sites, sites_b); H2D(d_lb, lb2, sizeof(double) * (size_t)B * n); H2D(d_ub, ub2, sizeof(double) * (size_t)B * n);
    H2D(d_ndb, n_db, sB); H2D(d_nf, n_found, sB); H2D(d_found, found, sizeof(int) * (size_t)B * found_stride);
    if (extra_stride > 0) { H2D(d_extra, extra_sites, extra_b); H2D(d_ne, n_extra, sB); }
    rc = run_round4(ctx, cfg, B, n, db_stride, d_sites, d_ndb, d_lb, d_ub, found_stride, d_found, d_nf, extra_stride,
                    extra_stride > 0 ? d_extra : nullptr, extra_stride > 0 ? d_ne : nullptr, found_stride + extra_stride,
                    r4_stride, d_r4, d_nr4, d_status);
    if (rc != MRBF_OK) return rc;
    D2H(r4, d_r4, sizeof(int) * (size_t)B * r4_stride); D2H(n_r4, d_nr4, sB);
    if (status) D2H(status, d_status, sB);
    CK(cudaStreamSynchronize(ctx->stream));
    return MRBF_OK;
}

int mrbf_gather_training_dev(mrbf_ctx* ctx, int32_t B, int32_t n, int32_t k, int32_t db_stride,
                             const double* sites, const double* values, const int32_t* x_index,
                             const int32_t* r1, const int32_t* n_r1, const int32_t* r2, const int32_t* n_r2,
                             const double* r3_sites, const double* r3_values, const int32_t* n_r3,
                             int32_t r4_stride, const int32_t* r4, const int32_t* n_r4,
                             int32_t train_stride, double* train_sites, double* train_values, int32_t* N) {
    if (!ctx) return MRBF_EINVAL;
    if (B <= 0 || n <= 0 || k <= 0) return fail(ctx, MRBF_EINVAL, "bad sizes%s");
    CK(cudaSetDevice(ctx->device));
    GatherParams G{};
    G.B = B; G.n = n; G.k = k; G.db_stride = db_stride; G.r4_stride = r4_stride; G.train_stride = train_stride;
    G.sites = sites; G.values = values; G.x_index = x_index; G.r1 = r1; G.n_r1 = n_r1; G.r2 = r2; G.n_r2 = n_r2;
    G.r3_sites = r3_sites; G.r3_values = r3_values; G.n_r3 = n_r3; G.r4 = r4; G.n_r4 = n_r4;
    G.train_sites = train_sites; G.train_values = train_values; G.N = N;
    { Timed t_(ctx, 2); CK(launch_gather_training(G, ctx->stream)); }
    ctx->launches += 1;
    return MRBF_OK;
}

// ------------------------------------------------------------------------------------------------ build
void mrbf_free_model(mrbf_ctx* ctx, mrbf_model* m) {
    if (!m) return;
    if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); }
    cudaFree(m->N); cudaFree(m->centers); cudaFree(m->w); cudaFree(m->lam); cudaFree(m->alpha2); cudaFree(m->pack);
    delete m;
}

static int build_impl(mrbf_ctx* ctx, const mrbf_cfg* cfg, int32_t B, int32_t n, int32_t k, int32_t train_stride,
                      const int32_t* N, const double* sites, const double* values, const double* shape,
                      mrbf_model** out, int32_t* status, const mrbf_prepared* kp, const double* db_values, const double* r3_values,
                      const int* skip_from_prepared, const double* db_sites = nullptr, const double* r3_sites = nullptr) {
    if (!ctx || !out) return MRBF_EINVAL;
    mrbf_model* recycle = *out;                 // NULL, or an earlier handle whose device buffers are reused when the shapes match
    *out = nullptr;
    // Ownership of the handle passed in moved into this call: every error return releases it (*model stays NULL).
    struct Drop { mrbf_ctx* c; mrbf_model** r; ~Drop() { if (*r) { mrbf_free_model(c, *r); *r = nullptr; } } } drop_{ctx, &recycle};
    int rc = check_cfg(ctx, cfg);
    if (rc != MRBF_OK) return rc;
    if (B <= 0 || n <= 0 || k <= 0 || k > 256 || train_stride <= 0) return fail(ctx, MRBF_EINVAL, "bad sizes%s");
    CK(cudaSetDevice(ctx->device));
    RadFn rf; double alpha; int cpd;
    rc = resolve_radfn(ctx, cfg, cfg->shape_parameter, &rf, &alpha, &cpd);
    if (rc != MRBF_OK) return rc;
    int deg = cfg->polynomial_degree;
    if (deg < cpd - 1) deg = cpd - 1;           // degree raised to cpd_order - 1 (assumption U4)
    if (deg > 2) return fail(ctx, MRBF_EUNSUPPORTED, "kernel needs a polynomial tail of degree > 2%s");
    const int p = poly_dim(n, deg), pl = p > 0 ? p : 1;
    mrbf_model* m = nullptr;
    cudaError_t e = cudaSuccess;
    if (recycle && recycle->B == B && recycle->n == n && recycle->k == k && recycle->train_stride == train_stride && recycle->p == p) {
        m = recycle; recycle = nullptr;         // same stream => ordered after every earlier use of the handle
        m->deg = deg; m->pack_valid = false;
    } else {
        if (recycle) { mrbf_free_model(ctx, recycle); recycle = nullptr; }
        m = new (std::nothrow) mrbf_model();
        if (!m) return fail(ctx, MRBF_ENOMEM, "out of host memory%s");
        m->B = B; m->n = n; m->k = k; m->train_stride = train_stride; m->p = p; m->deg = deg;
        if (e == cudaSuccess) e = dev_malloc(&m->N, sizeof(int) * (size_t)B);
        if (e == cudaSuccess) e = dev_malloc(&m->centers, sizeof(double) * (size_t)B * train_stride * n);
        if (e == cudaSuccess) e = dev_malloc(&m->w, sizeof(double) * (size_t)B * train_stride * k);
        if (e == cudaSuccess) e = dev_malloc(&m->lam, sizeof(double) * (size_t)B * pl * k);
        if (e == cudaSuccess) e = dev_malloc(&m->alpha2, sizeof(double) * (size_t)B);
        if (e == cudaSuccess && n <= 64 && k <= 16 && deg <= 1) {   // geometry of the tiled copy; the buffer itself is allocated on first use
            m->pack_s = eval_pack_stride(n); m->pack_nt = (train_stride + 63) / 64;
            m->pack_tile_doubles = (size_t)64 * m->pack_s + 64 + (size_t)k * 64;
        }
        if (e != cudaSuccess) { (void)cudaGetLastError(); mrbf_free_model(ctx, m); return fail(ctx, MRBF_ENOMEM, "cudaMalloc failed: %s", cudaGetErrorString(e)); }
    }
    m->kernel = rf.kernel; m->ibeta = rf.ibeta; m->sgn = rf.sgn;
    BuildParams Pb{};
    Pb.B = B; Pb.n = n; Pb.k = k; Pb.train_stride = train_stride; Pb.p = p; Pb.deg = deg;
    Pb.kernel = rf.kernel; Pb.ibeta = rf.ibeta; Pb.sgn = rf.sgn; Pb.alpha_default = alpha;
    Pb.N = N; Pb.sites = sites; Pb.values = values; Pb.shape = shape;
    if (kp) { Pb.skip = skip_from_prepared; Pb.centers_out = m->centers; Pb.N_out = m->N; }
    Pb.w = m->w; Pb.lam = m->lam; Pb.alpha2_out = m->alpha2; Pb.status = status; Pb.ld = train_stride | 1;
    const size_t vecd = build_vec_doubles(n, k, Pb.ld, p), wsd = build_ws_doubles(n, k, Pb.ld, p);
    size_t smem = vecd * sizeof(double);
    if ((vecd + wsd) * sizeof(double) <= SMEM_LIMIT) {
        Pb.ws_in_smem = 1; smem = (vecd + wsd) * sizeof(double); Pb.smem_ws_doubles = (int)wsd;
        // room left: stage the sites coordinate-major for the Gram-matrix assembly (row-per-lane reads of the AoS sites in global
        // memory cost 32 L1 wavefronts per load instruction -- 40 % of the LSU traffic of the whole kernel)
        const size_t stage = (size_t)n * (size_t)(train_stride | 1);
        // only when the system alone already limits the SM to one CTA: for small systems the extra shared memory costs residency
        if (smem > 113 * 1024 && (vecd + wsd + stage) * sizeof(double) <= SMEM_LIMIT) { Pb.stage_off = (int)(vecd + wsd); smem += stage * sizeof(double); }
    }
    else {
        // worst case does not fit: give every CTA the whole shared memory; instances whose own N fits use it,
        // the rest fall back to the global workspace
        if (smem > SMEM_LIMIT) { mrbf_free_model(ctx, m); return fail(ctx, MRBF_EUNSUPPORTED, "training set too large%s"); }
        Pb.ws_in_smem = 0; Pb.ws_stride = wsd;
        smem = SMEM_LIMIT;
        Pb.smem_ws_doubles = (int)(SMEM_LIMIT / sizeof(double) - vecd);
        int r_ = ensure(ctx, ctx->ws[8], (size_t)B * wsd * sizeof(double));
        if (r_ != MRBF_OK) { mrbf_free_model(ctx, m); return r_; }
        Pb.ws = (double*)ctx->ws[8].p;
    }
    e = cudaSuccess;
    if (!kp) {
        e = cudaMemcpyAsync(m->N, N, sizeof(int) * (size_t)B, cudaMemcpyDeviceToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(m->centers, sites, sizeof(double) * (size_t)B * train_stride * n, cudaMemcpyDeviceToDevice, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaMemsetAsync(m->w, 0, sizeof(double) * (size_t)B * train_stride * k, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->lam, 0, sizeof(double) * (size_t)B * pl * k, ctx->stream);
    // 0. no kept factorisation: the reduced-system route on the register-tiled round-4 kernels in BUILD MODE (found set = the first p
    //    training points, every other point a candidate that must be accepted: Pi_0^{-1}, panels, blocked Cholesky of N' Phi N in
    //    registers, then the two triangular solves of build_schur_kernel).  Instances that do not qualify (N <= p, more than 128
    //    reduced unknowns, ill-conditioned first p points, reduced matrix not positive definite) are left to the general kernel.
    // (the geometry covers at most 128 reduced unknowns; an instance with more falls back on its own)
    const SchurGeom bgeom = (p > 0 && deg <= 1 && train_stride > p) ? round4_schur_geom(n, p, train_stride - p < 128 ? train_stride - p : 128) : SchurGeom{};
    const char* bg_env = getenv("MRBF_BUILD_GENERAL");
    const bool reduced_route = !kp && e == cudaSuccess && bgeom.eligible && !(bg_env && atoi(bg_env) != 0) &&
                               build_schur_smem_doubles(k, bgeom.MC, p) * sizeof(double) <= SMEM_LIMIT;
    if (reduced_route) {
        const size_t ni = (size_t)B * 4 + (size_t)B * p + (size_t)B * bgeom.MC;
        int r_ = ensure(ctx, ctx->ws[16], (size_t)B * bgeom.state_doubles * sizeof(double));
        if (r_ == MRBF_OK) r_ = ensure(ctx, ctx->ws[17], ni * sizeof(int));
        if (r_ == MRBF_OK) r_ = ensure(ctx, ctx->ws[13], (size_t)B * bgeom.pw_doubles * sizeof(double));
        if (r_ != MRBF_OK) { mrbf_free_model(ctx, m); return r_; }
        int* b_elig = (int*)ctx->ws[17].p; int* b_nfound = b_elig + B; int* b_nr4 = b_nfound + B; int* b_done = b_nr4 + B;
        int* b_found = b_done + B; int* b_r4 = b_found + (size_t)B * p;
        Round4Params R{};
        R.B = B; R.n = n; R.db_stride = train_stride; R.found_stride = p; R.extra_stride = 0; R.r4_stride = bgeom.MC;
        R.NM = train_stride; R.max_points = 1 << 30;
        R.cfg.polynomial_degree = deg; R.cfg.optimized_sampling = 1;
        R.rf = rf; R.chol_thr = 0.0;
        R.sites = sites; R.n_db = N; R.r4 = b_r4; R.n_r4 = b_nr4; R.status = nullptr;
        R.keep_fs = (double*)ctx->ws[16].p; R.fs_stride = bgeom.state_doubles; R.elig = b_elig;
        R.panel_ws = (double*)ctx->ws[13].p;
        R.build_mode = 1; R.shape_arr = shape; R.alpha_default = alpha; R.alpha2_out = m->alpha2;
        R.found_out = b_found; R.n_found_out = b_nfound;
        e = cudaMemsetAsync(b_done, 0, sizeof(int) * (size_t)B, ctx->stream);
        if (e == cudaSuccess) { Timed t_(ctx, 1); e = launch_round4_schur(R, bgeom, ctx->stream); ctx->launches += 1; }
        if (e == cudaSuccess) {
            SchurBuildParams Q{};
            Q.B = B; Q.n = n; Q.k = k; Q.p = p; Q.deg = deg; Q.db_stride = train_stride; Q.found_stride = p;
            Q.r4_stride = bgeom.MC; Q.train_stride = train_stride; Q.MC = bgeom.MC; Q.fs_stride = bgeom.state_doubles;
            Q.off_M0 = bgeom.off_M0; Q.off_U = bgeom.off_U; Q.off_C = bgeom.off_C; Q.off_L = bgeom.off_L; Q.off_acc = bgeom.off_acc;
            Q.fs = R.keep_fs; Q.elig = b_elig; Q.found = b_found; Q.n_found = b_nfound; Q.r4 = b_r4;
            Q.sites = sites; Q.values = values; Q.r3_sites = nullptr; Q.r3_values = nullptr; Q.alpha2 = -1.0;
            Q.centers = m->centers; Q.w = m->w; Q.lam = m->lam; Q.alpha2_out = m->alpha2; Q.N = m->N; Q.status = status; Q.done = b_done;
            Timed t_(ctx, 6);
            e = launch_build_schur(Q, build_schur_smem_doubles(k, bgeom.MC, p) * sizeof(double), ctx->stream);
            ctx->launches += 1;
            Pb.skip = b_done;
        }
    }
    if (e == cudaSuccess && kp && kp->kind == 1 && kp->cfg_degree == deg && kp->p > 0) {
        // 1'. factorisation kept by round4_elim_kernel: two triangular solves per output
        SchurBuildParams Q{};
        const SchurGeom& g = kp->geom;
        Q.B = B; Q.n = n; Q.k = k; Q.p = kp->p; Q.deg = deg; Q.db_stride = kp->db_stride; Q.found_stride = kp->found_stride;
        Q.r4_stride = kp->r4_stride; Q.train_stride = train_stride; Q.MC = g.MC; Q.fs_stride = kp->fs_stride;
        Q.off_M0 = g.off_M0; Q.off_U = g.off_U; Q.off_C = g.off_C; Q.off_L = g.off_L; Q.off_acc = g.off_acc;
        Q.fs = kp->fs; Q.elig = kp->elig; Q.found = kp->found; Q.n_found = kp->n_found; Q.r4 = kp->r4;
        Q.sites = db_sites; Q.values = db_values; Q.r3_sites = r3_sites; Q.r3_values = r3_values; Q.alpha2 = alpha * alpha;
        Q.centers = m->centers; Q.w = m->w; Q.lam = m->lam; Q.alpha2_out = m->alpha2; Q.N = m->N; Q.status = status; Q.done = (int*)skip_from_prepared;
        const size_t psm = build_schur_smem_doubles(k, g.MC, kp->p) * sizeof(double);
        if (psm <= SMEM_LIMIT) { Timed t_(ctx, 6); e = launch_build_schur(Q, psm, ctx->stream); ctx->launches += 1; }
        else e = cudaMemsetAsync((void*)skip_from_prepared, 0, sizeof(int) * (size_t)B, ctx->stream);
    } else if (e == cudaSuccess && kp && kp->kind == 0 && kp->cfg_degree == deg && kp->p > 0) {
        // 1. instances whose round 4 kept its factorisation: two triangular mat-vecs per output
        PreparedBuildParams Q{};
        Q.B = B; Q.n = n; Q.k = k; Q.NM = kp->NM; Q.p = kp->p; Q.deg = deg; Q.db_stride = kp->db_stride;
        Q.found_stride = kp->found_stride; Q.r4_stride = kp->r4_stride; Q.train_stride = train_stride; Q.fs_stride = kp->fs_stride;
        round4_fast_state_layout(n, kp->NM, kp->p, &Q.off_M0, &Q.off_G, &Q.off_C, &Q.off_L);
        Q.fs = kp->fs; Q.elig = kp->elig; Q.found = kp->found; Q.n_found = kp->n_found; Q.n_extra = kp->n_extra; Q.r4 = kp->r4; Q.n_r4 = kp->n_r4;
        Q.values = db_values; Q.r3_values = r3_values; Q.alpha2 = alpha * alpha;
        Q.centers = m->centers; Q.w = m->w; Q.lam = m->lam; Q.alpha2_out = m->alpha2; Q.N = m->N; Q.status = status; Q.done = (int*)skip_from_prepared;
        const size_t psm = build_prepared_smem_doubles(n, k, kp->NM, kp->p) * sizeof(double);
        const size_t ssm = build_prepared_stream_smem_doubles(n, k, kp->NM, kp->p) * sizeof(double);
        if (psm <= SMEM_LIMIT) { Timed t_(ctx, 6); e = launch_build_prepared(Q, psm, ctx->stream); ctx->launches += 1; }
        else if (ssm <= SMEM_LIMIT) { Timed t_(ctx, 6); e = launch_build_prepared_stream(Q, ssm, ctx->stream); ctx->launches += 1; }   // L^-1 streamed
        else e = cudaMemsetAsync((void*)skip_from_prepared, 0, sizeof(int) * (size_t)B, ctx->stream);
    } else if (e == cudaSuccess && kp) {
        e = cudaMemsetAsync((void*)skip_from_prepared, 0, sizeof(int) * (size_t)B, ctx->stream);
    }
    // 2. general route (for every instance, or for the ones step 1 left)
    if (e == cudaSuccess) { Timed t_(ctx, 3); e = launch_build(Pb, smem, ctx->stream); }
    if (e != cudaSuccess) { mrbf_free_model(ctx, m); return fail(ctx, MRBF_ECUDA, "CUDA error: %s", cudaGetErrorString(e)); }
    ctx->launches += 1;
    *out = m;
    return MRBF_OK;
}

static int mrbf_build_dev_inner(mrbf_ctx* ctx, const mrbf_cfg* cfg, int32_t B, int32_t n, int32_t k, int32_t train_stride,
                   const int32_t* N, const double* sites, const double* values, const double* shape,
                   mrbf_model** out, int32_t* status) {
    return build_impl(ctx, cfg, B, n, k, train_stride, N, sites, values, shape, out, status, nullptr, nullptr, nullptr, nullptr);
}

static int mrbf_build_prepared_dev_inner(mrbf_ctx* ctx, const mrbf_cfg* cfg, const mrbf_prepared* kp, int32_t k,
                            const double* sites, const double* values, const double* r3_sites, const double* r3_values,
                            const int32_t* x_index, const int32_t* r1, const int32_t* n_r1, const int32_t* r2, const int32_t* n_r2,
                            const int32_t* n_r3, mrbf_model** out, int32_t* status) {
    if (!ctx || !kp || !out || !cfg) return MRBF_EINVAL;
    if (cfg->kernel != kp->kernel || !((cfg->shape_parameter == kp->shape) || (cfg->shape_parameter != cfg->shape_parameter && kp->shape != kp->shape)))
        return fail(ctx, MRBF_EINVAL, "mrbf_build_prepared: kernel/shape differ from the ones the factorisation was made with%s");
    CK(cudaSetDevice(ctx->device));
    const int B = kp->B, n = kp->n, ts = kp->NM;
    // scratch for the general route: gathered training sets + done mask
    ENSURE(ctx->ws[10], sizeof(double) * (size_t)B * ts * (n + k));
    ENSURE(ctx->ws[11], sizeof(int) * (size_t)B * 2);
    double* tsit = (double*)ctx->ws[10].p; double* tval = tsit + (size_t)B * ts * n;
    int* done = (int*)ctx->ws[11].p; int* Ntmp = done + B;
    CK(cudaMemsetAsync(done, 0, sizeof(int) * (size_t)B, ctx->stream));
    // the general route needs the gathered training set of the instances the kept factorisation does not cover;
    // the gather is cheap, so it runs for all and the build kernel skips the done ones
    GatherParams G{};
    G.B = B; G.n = n; G.k = k; G.db_stride = kp->db_stride; G.r4_stride = kp->r4_stride; G.train_stride = ts;
    G.sites = sites; G.values = values; G.x_index = x_index; G.r1 = r1; G.n_r1 = n_r1; G.r2 = r2; G.n_r2 = n_r2;
    G.r3_sites = r3_sites; G.r3_values = r3_values; G.n_r3 = n_r3; G.r4 = kp->r4; G.n_r4 = kp->n_r4;
    G.train_sites = tsit; G.train_values = tval; G.N = Ntmp;
    {   // instances whose factorisation was kept (and will be used: same conditions as in build_impl) never take the general route
        RadFn rf_; double a_; int cpd_ = 0;
        int deg_ = cfg->polynomial_degree;
        if (resolve_radfn(ctx, cfg, cfg->shape_parameter, &rf_, &a_, &cpd_) == MRBF_OK && deg_ < cpd_ - 1) deg_ = cpd_ - 1;
        const bool schur_build = kp->kind == 1 && kp->cfg_degree == deg_ && kp->p > 0 &&
                                 build_schur_smem_doubles(k, kp->geom.MC, kp->p) * sizeof(double) <= SMEM_LIMIT;
        G.skip = schur_build ? kp->elig : nullptr;
    }
    { Timed t_(ctx, 2); CK(launch_gather_training(G, ctx->stream)); }
    ctx->launches += 1;
    return build_impl(ctx, cfg, B, n, k, ts, Ntmp, tsit, tval, nullptr, out, status, kp, values, r3_values, done, sites, r3_sites);
}

static int mrbf_build_prepared_inner(mrbf_ctx* ctx, const mrbf_cfg* cfg, const mrbf_prepared* kp, int32_t k,
                        const double* sites, const double* values, const double* r3_sites, const double* r3_values,
                        const int32_t* x_index, const int32_t* r1, const int32_t* n_r1, const int32_t* r2, const int32_t* n_r2,
                        const int32_t* n_r3, mrbf_model** out, int32_t* status) {
    if (!ctx || !kp || !out || !cfg || !sites || !values || !x_index || !r1 || !n_r1 || !r2 || !n_r2 || !n_r3) return MRBF_EINVAL;
    if (k <= 0) return fail(ctx, MRBF_EINVAL, "bad sizes%s");
    CK(cudaSetDevice(ctx->device));
    const int B = kp->B, n = kp->n, dbs = kp->db_stride;
    const size_t sb = sizeof(double) * (size_t)B * dbs * n, vb = sizeof(double) * (size_t)B * dbs * k;
    const size_t r3s = sizeof(double) * (size_t)B * n * n, r3v = sizeof(double) * (size_t)B * n * k, sB = sizeof(int) * (size_t)B;
    ENSURE(ctx->hb[21], sb + vb + r3s + r3v); ENSURE(ctx->hb[22], sB * 5 + 2 * sizeof(int) * (size_t)B * n);
    double* d_s = (double*)ctx->hb[21].p; double* d_v = d_s + (size_t)B * dbs * n; double* d_r3s = d_v + (size_t)B * dbs * k;
    double* d_r3v = d_r3s + (size_t)B * n * n;
    int* d_xi = (int*)ctx->hb[22].p; int* d_n1 = d_xi + B; int* d_n2 = d_n1 + B; int* d_n3 = d_n2 + B; int* d_st = d_n3 + B;
    int* d_r1 = d_st + B; int* d_r2 = d_r1 + (size_t)B * n;
    H2D(d_s, sites, sb); H2D(d_v, values, vb);
    if (r3_sites) H2D(d_r3s, r3_sites, r3s); else CK(cudaMemsetAsync(d_r3s, 0, r3s, ctx->stream));
    if (r3_values) H2D(d_r3v, r3_values, r3v); else CK(cudaMemsetAsync(d_r3v, 0, r3v, ctx->stream));
    H2D(d_xi, x_index, sB); H2D(d_n1, n_r1, sB); H2D(d_n2, n_r2, sB); H2D(d_n3, n_r3, sB);
    H2D(d_r1, r1, sizeof(int) * (size_t)B * n); H2D(d_r2, r2, sizeof(int) * (size_t)B * n);
    int rc = mrbf_build_prepared_dev_inner(ctx, cfg, kp, k, d_s, d_v, d_r3s, d_r3v, d_xi, d_r1, d_n1, d_r2, d_n2, d_n3, out, d_st);
    if (rc != MRBF_OK) return rc;
    if (status) D2H(status, d_st, sB);
    CK(cudaStreamSynchronize(ctx->stream));
    if (status) for (int b = 0; b < B; ++b) if (status[b] != 0) return fail(ctx, MRBF_ENUMERIC, "at least one system failed; see status[]%s");
    return MRBF_OK;
}

static int mrbf_build_inner(mrbf_ctx* ctx, const mrbf_cfg* cfg, int32_t B, int32_t n, int32_t k, int32_t train_stride,
               const int32_t* N, const double* sites, const double* values, const double* shape,
               mrbf_model** out, int32_t* status) {
    if (!ctx || !out) return MRBF_EINVAL;
    if (B <= 0 || n <= 0 || k <= 0 || train_stride <= 0) return fail(ctx, MRBF_EINVAL, "bad sizes%s");
    CK(cudaSetDevice(ctx->device));
    const size_t sb = sizeof(double) * (size_t)B * train_stride * n, vb = sizeof(double) * (size_t)B * train_stride * k;
    ENSURE(ctx->hb[8], sb); ENSURE(ctx->hb[9], vb); ENSURE(ctx->hb[10], sizeof(int) * (size_t)B * 2 + sizeof(double) * (size_t)B);
    double* d_s = (double*)ctx->hb[8].p; double* d_v = (double*)ctx->hb[9].p;
    double* d_shape = (double*)ctx->hb[10].p; int* d_N = (int*)(d_shape + B); int* d_status = d_N + B;
    H2D(d_s, sites, sb); H2D(d_v, values, vb); H2D(d_N, N, sizeof(int) * (size_t)B);
    if (shape) H2D(d_shape, shape, sizeof(double) * (size_t)B);
    int rc = mrbf_build_dev_inner(ctx, cfg, B, n, k, train_stride, d_N, d_s, d_v, shape ? d_shape : nullptr, out, d_status);
    if (rc != MRBF_OK) return rc;
    if (status) D2H(status, d_status, sizeof(int) * (size_t)B);
    CK(cudaStreamSynchronize(ctx->stream));
    if (status) for (int b = 0; b < B; ++b) if (status[b] != 0) return fail(ctx, MRBF_ENUMERIC, "at least one system failed; see status[]%s");
    return MRBF_OK;
}

// Public wrappers: one ownership rule for the in/out model handle.  MRBF_OK and MRBF_ENUMERIC (some instances failed
// numerically, see status[]) return a valid handle; every other error releases whatever handle is involved and returns NULL.
static int settle_model(mrbf_ctx* ctx, mrbf_model** out, int rc) {
    if (rc != MRBF_OK && rc != MRBF_ENUMERIC && out && *out) { mrbf_free_model(ctx, *out); *out = nullptr; }
    return rc;
}
int mrbf_build_dev(mrbf_ctx* ctx, const mrbf_cfg* cfg, int32_t B, int32_t n, int32_t k, int32_t train_stride,
                   const int32_t* N, const double* sites, const double* values, const double* shape,
                   mrbf_model** out, int32_t* status) {
    return settle_model(ctx, out, mrbf_build_dev_inner(ctx, cfg, B, n, k, train_stride, N, sites, values, shape, out, status));
}
int mrbf_build(mrbf_ctx* ctx, const mrbf_cfg* cfg, int32_t B, int32_t n, int32_t k, int32_t train_stride,
               const int32_t* N, const double* sites, const double* values, const double* shape,
               mrbf_model** out, int32_t* status) {
    return settle_model(ctx, out, mrbf_build_inner(ctx, cfg, B, n, k, train_stride, N, sites, values, shape, out, status));
}
int mrbf_build_prepared_dev(mrbf_ctx* ctx, const mrbf_cfg* cfg, const mrbf_prepared* kp, int32_t k,
                            const double* sites, const double* values, const double* r3_sites, const double* r3_values,
                            const int32_t* x_index, const int32_t* r1, const int32_t* n_r1, const int32_t* r2, const int32_t* n_r2,
                            const int32_t* n_r3, mrbf_model** out, int32_t* status) {
    return settle_model(ctx, out, mrbf_build_prepared_dev_inner(ctx, cfg, kp, k, sites, values, r3_sites, r3_values, x_index, r1, n_r1, r2, n_r2,
                                                                n_r3, out, status));
}
int mrbf_build_prepared(mrbf_ctx* ctx, const mrbf_cfg* cfg, const mrbf_prepared* kp, int32_t k,
                        const double* sites, const double* values, const double* r3_sites, const double* r3_values,
                        const int32_t* x_index, const int32_t* r1, const int32_t* n_r1, const int32_t* r2, const int32_t* n_r2,
                        const int32_t* n_r3, mrbf_model** out, int32_t* status) {
    return settle_model(ctx, out, mrbf_build_prepared_inner(ctx, cfg, kp, k, sites, values, r3_sites, r3_values, x_index, r1, n_r1, r2, n_r2,
                                                            n_r3, out, status));
}

int mrbf_model_dims(const mrbf_model* m, int32_t* out6) {
    if (!m || !out6) return MRBF_EINVAL;
    out6[0] = m->B; out6[1] = m->n; out6[2] = m->k; out6[3] = m->train_stride; out6[4] = m->p; out6[5] = m->deg;
    return MRBF_OK;
}

int mrbf_model_coeffs(mrbf_ctx* ctx, const mrbf_model* m, double* w, double* lambda) {
    if (!ctx || !m) return MRBF_EINVAL;
    CK(cudaSetDevice(ctx->device));
    if (w) D2H(w, m->w, sizeof(double) * (size_t)m->B * m->train_stride * m->k);
    if (lambda && m->p > 0) D2H(lambda, m->lam, sizeof(double) * (size_t)m->B * m->p * m->k);
    CK(cudaStreamSynchronize(ctx->stream));
    return MRBF_OK;
}

// ------------------------------------------------------------------------------------------------ eval
static void fill_eval(EvalParams& E, const mrbf_model* m, int64_t M, const double* X, double* Y, double* J) {
    E.B = m->B; E.n = m->n; E.k = m->k; E.train_stride = m->train_stride; E.p = m->p; E.deg = m->deg;
    E.kernel = m->kernel; E.ibeta = m->ibeta; E.sgn = m->sgn; E.M = M;
    E.N = m->N; E.centers = m->centers; E.w = m->w; E.lam = m->lam; E.alpha2 = m->alpha2; E.X = X; E.Y = Y; E.J = J;
    E.pack = m->pack_valid ? m->pack : nullptr; E.pack_s = m->pack_s; E.pack_nt = m->pack_nt; E.pack_tile_doubles = m->pack_tile_doubles;
}

int mrbf_eval_dev(mrbf_ctx* ctx, const mrbf_model* m, int64_t M, const double* X, double* Y, double* J) {
    if (!ctx || !m || !X || M < 0) return MRBF_EINVAL;
    if (M == 0 || (!Y && !J)) return MRBF_OK;
    CK(cudaSetDevice(ctx->device));
    int nl = 0;
    if (m->pack_tile_doubles && !m->pack_valid && M > 8) {   // (a handful of points goes through eval_small_kernel, which needs no tiles)
        // first evaluation of this model: re-tile it once for the tensor-path sweep (the handle is logically const)
        mrbf_model* mm = const_cast<mrbf_model*>(m);
        if (!mm->pack) {
            cudaError_t e = dev_malloc(&mm->pack, sizeof(double) * (size_t)m->B * m->pack_nt * m->pack_tile_doubles);
            if (e != cudaSuccess) { (void)cudaGetLastError(); mm->pack = nullptr; mm->pack_tile_doubles = 0; }   // handled: fall back to the tile kernels
        }
        if (mm->pack) {
            PackParams K{};
            K.B = m->B; K.n = m->n; K.k = m->k; K.train_stride = m->train_stride; K.s = m->pack_s; K.nt = m->pack_nt;
            K.tile_doubles = m->pack_tile_doubles; K.N = m->N; K.centers = m->centers; K.w = m->w; K.pack = mm->pack;
            CK(launch_eval_pack(K, ctx->stream));
            ctx->launches += 1;
            mm->pack_valid = true;
        }
    }
    EvalParams E{};
    fill_eval(E, m, M, X, Y, J);       // with J the same pass also produces the values
    if (E.pack && M > 8 && m->n <= 64 && m->B <= 65535 && m->k <= 16) {
        // grids below one wave of the tensor-path sweep (Armijo / PS batches of a single run, the low end of the C5 sweep): split the
        // centre tiles over several CTAs per point tile; partial sums are added in a fixed order by eval_split_reduce_kernel
        const int z = eval_split_factor(M, m->B, m->pack_nt);
        if (z > 1) {
            const size_t ny = (size_t)m->B * M * m->k, nj = J ? ny * m->n : 0;
            ENSURE(ctx->ws[18], sizeof(double) * (size_t)z * (ny + nj));
            E.zsplit = z; E.partY = (double*)ctx->ws[18].p; E.partJ = J ? E.partY + (size_t)z * ny : nullptr;
        }
    }
    { Timed t_(ctx, 4); CK(launch_eval(E, ctx->stream, &nl)); }
    ctx->launches += nl;
    return MRBF_OK;
}

int mrbf_eval(mrbf_ctx* ctx, const mrbf_model* m, int64_t M, const double* X, double* Y, double* J) {
    if (!ctx || !m || !X || M < 0) return MRBF_EINVAL;
    if (M == 0 || (!Y && !J)) return MRBF_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t xb = sizeof(double) * (size_t)m->B * M * m->n, yb = sizeof(double) * (size_t)m->B * M * m->k;
    const size_t jb = yb * m->n;
    ENSURE(ctx->hb[11], xb);
    if (Y) ENSURE(ctx->hb[12], yb);
    if (J) ENSURE(ctx->hb[13], jb);
    double* dX = (double*)ctx->hb[11].p; double* dY = Y ? (double*)ctx->hb[12].p : nullptr; double* dJ = J ? (double*)ctx->hb[13].p : nullptr;
    H2D(dX, X, xb);
    int rc = mrbf_eval_dev(ctx, m, M, dX, dY, dJ);
    if (rc != MRBF_OK) return rc;
    if (Y) D2H(Y, dY, yb);
    if (J) D2H(J, dJ, jb);
    CK(cudaStreamSynchronize(ctx->stream));
    return MRBF_OK;
}

int mrbf_backtrack_dev(mrbf_ctx* ctx, const mrbf_model* m, const double* x, const double* dir, const double* step0,
                       const double* omega, double armijo_c, double shrink, double min_stepsize, int32_t max_loops,
                       int32_t strict, int32_t* step_index, double* sigma, double* x_plus, double* mx, double* mx_plus) {
    if (!ctx || !m || !x || !dir || !step0 || !omega || max_loops < 0 || !step_index || !sigma || !x_plus || !mx || !mx_plus) return MRBF_EINVAL;
    CK(cudaSetDevice(ctx->device));
    const int B = m->B, n = m->n, k = m->k, ns = max_loops + 1;
    ENSURE(ctx->hb[15], sizeof(double) * (size_t)B * (ns + 1) * (n + k) + sizeof(double) * (size_t)B * ns);
    double* d_Xall = (double*)ctx->hb[15].p; double* d_Yall = d_Xall + (size_t)B * (ns + 1) * n; double* d_sall = d_Yall + (size_t)B * (ns + 1) * k;
    BacktrackParams Q{};
    Q.B = B; Q.n = n; Q.k = k; Q.nsteps = ns; Q.strict = strict; Q.armijo_c = armijo_c; Q.shrink = shrink;
    Q.min_stepsize = min_stepsize; Q.max_loops = max_loops;
    Q.x = x; Q.dir = dir; Q.step0 = step0; Q.omega = omega; Q.Yall = d_Yall; Q.Xall = d_Xall; Q.sig_all = d_sall;
    Q.step_index = step_index; Q.sigma = sigma; Q.x_plus = x_plus; Q.mx = mx; Q.mx_plus = mx_plus;
    CK(launch_backtrack_points(Q, ctx->stream));
    ctx->launches += 1;
    int rc = mrbf_eval_dev(ctx, m, ns + 1, d_Xall, d_Yall, nullptr);
    if (rc != MRBF_OK) return rc;
    CK(launch_backtrack_pick(Q, ctx->stream));
    ctx->launches += 1;
    return MRBF_OK;
}

int mrbf_backtrack(mrbf_ctx* ctx, const mrbf_model* m, const double* x, const double* dir, const double* step0,
                   const double* omega, double armijo_c, double shrink, double min_stepsize, int32_t max_loops,
                   int32_t strict, int32_t* step_index, double* sigma, double* x_plus, double* mx, double* mx_plus) {
    if (!ctx || !m || !x || !dir || !step0 || !omega || max_loops < 0) return MRBF_EINVAL;
    CK(cudaSetDevice(ctx->device));
    const int B = m->B, n = m->n, k = m->k;
    const size_t dBn = sizeof(double) * (size_t)B * n, dB = sizeof(double) * (size_t)B, dBk = sizeof(double) * (size_t)B * k;
    ENSURE(ctx->hb[14], 3 * dBn + 3 * dB + 2 * dBk + sizeof(int) * (size_t)B);
    double* d_x = (double*)ctx->hb[14].p; double* d_dir = d_x + (size_t)B * n; double* d_xp = d_dir + (size_t)B * n;
    double* d_step = d_xp + (size_t)B * n; double* d_om = d_step + B; double* d_sig = d_om + B;
    double* d_mx = d_sig + B; double* d_mxp = d_mx + (size_t)B * k; int* d_idx = (int*)(d_mxp + (size_t)B * k);
    H2D(d_x, x, dBn); H2D(d_dir, dir, dBn); H2D(d_step, step0, dB); H2D(d_om, omega, dB);
    int rc = mrbf_backtrack_dev(ctx, m, d_x, d_dir, d_step, d_om, armijo_c, shrink, min_stepsize, max_loops, strict, d_idx, d_sig, d_xp, d_mx, d_mxp);
    if (rc != MRBF_OK) return rc;
    if (step_index) D2H(step_index, d_idx, sizeof(int) * (size_t)B);
    if (sigma) D2H(sigma, d_sig, dB);
    if (x_plus) D2H(x_plus, d_xp, dBn);
    if (mx) D2H(mx, d_mx, dBk);
    if (mx_plus) D2H(mx_plus, d_mxp, dBk);
    CK(cudaStreamSynchronize(ctx->stream));
    return MRBF_OK;
}


// ------------------------------------------------------------------------------------------------ steepest-descent LP
int mrbf_descent_direction_dev(mrbf_ctx* ctx, int32_t B, int32_t n, int32_t k, const double* jac, const double* x,
                               const double* lb, const double* ub, int32_t normalize,
                               double* d, double* omega, int32_t* iters, int32_t* status) {
    if (!ctx || !jac || !x || !lb || !ub || !d || !omega) return MRBF_EINVAL;
    if (B <= 0 || n <= 0 || k <= 0) return fail(ctx, MRBF_EINVAL, "bad sizes%s");
    if (k > descent_max_outputs()) return fail(ctx, MRBF_EUNSUPPORTED, "mrbf_descent_direction supports at most 8 outputs%s");
    CK(cudaSetDevice(ctx->device));
    DescentParams P{};
    P.B = B; P.n = n; P.k = k; P.normalize = normalize ? 1 : 0; P.warp_doubles = descent_warp_doubles(n, k);
    if (4 * P.warp_doubles * sizeof(double) > SMEM_LIMIT) return fail(ctx, MRBF_EUNSUPPORTED, "too many variables for mrbf_descent_direction%s");
    P.jac = jac; P.x = x; P.lb = lb; P.ub = ub; P.d = d; P.omega = omega; P.iters = iters; P.status = status;
    CK(launch_descent_direction(P, ctx->stream));
    ctx->launches += 1;
    return MRBF_OK;
}

int mrbf_descent_direction(mrbf_ctx* ctx, int32_t B, int32_t n, int32_t k, const double* jac, const double* x,
                           const double* lb, const double* ub, int32_t normalize,
                           double* d, double* omega, int32_t* iters, int32_t* status) {
    if (!ctx || !jac || !x || !lb || !ub || !d || !omega) return MRBF_EINVAL;
    if (B <= 0 || n <= 0 || k <= 0) return fail(ctx, MRBF_EINVAL, "bad sizes%s");
    CK(cudaSetDevice(ctx->device));
    const size_t bj = sizeof(double) * (size_t)B * k * n, bx = sizeof(double) * (size_t)B * n, bn = sizeof(double) * (size_t)n;
    ENSURE(ctx->hb[20], bj + 2 * bx + 2 * bn + sizeof(double) * (size_t)B + 2 * sizeof(int) * (size_t)B);
    double* d_j = (double*)ctx->hb[20].p; double* d_x = d_j + (size_t)B * k * n; double* d_d = d_x + (size_t)B * n;
    double* d_lb = d_d + (size_t)B * n; double* d_ub = d_lb + n; double* d_om = d_ub + n;
    int* d_it = (int*)(d_om + B); int* d_st = d_it + B;
    H2D(d_j, jac, bj); H2D(d_x, x, bx); H2D(d_lb, lb, bn); H2D(d_ub, ub, bn);
    int rc = mrbf_descent_direction_dev(ctx, B, n, k, d_j, d_x, d_lb, d_ub, normalize, d_d, d_om, d_it, d_st);
    if (rc != MRBF_OK) return rc;
    D2H(d, d_d, bx); D2H(omega, d_om, sizeof(double) * (size_t)B);
    if (iters) D2H(iters, d_it, sizeof(int) * (size_t)B);
    if (status) D2H(status, d_st, sizeof(int) * (size_t)B);
    CK(cudaStreamSynchronize(ctx->stream));
    return MRBF_OK;
}

// ------------------------------------------------------------------------------------------------ Pascoletti-Serafini inner solves
int mrbf_ps_solve_dev(mrbf_ctx* ctx, const mrbf_model* m, const double* x, const double* lb, const double* ub,
                      const double* mx, const double* dir, int32_t n_obj, int32_t objective, int32_t population,
                      int32_t max_evals, int64_t seed, double* f_min, double* x_min, double* y_min, int32_t* found,
                      int32_t* evals_used) {
    if (!ctx || !m || !x || !lb || !ub || !f_min || !x_min || !y_min || !found) return MRBF_EINVAL;
    const int B = m->B, n = m->n, k = m->k;
    if (n_obj <= 0 || n_obj > k) return fail(ctx, MRBF_EINVAL, "n_obj must lie in 1..k%s");
    if (dir && !mx) return fail(ctx, MRBF_EINVAL, "Pascoletti-Serafini mode needs m(x)%s");
    if (!dir && (objective < 0 || objective >= n_obj)) return fail(ctx, MRBF_EINVAL, "ideal-point mode needs an objective index in 0..n_obj-1%s");
    int lam = population > 0 ? population : 20 * (n + 1);              // NLopt's ISRES default: 20 (n + 1)
    if (lam < 8) lam = 8;
    const int evals = max_evals > 0 ? max_evals : 500 * (n + 1);        // descent.jl:373, 418, 535
    int gens = evals / lam; if (gens < 1) gens = 1;                     // generations after the initial population
    if (ps_rank_smem_bytes(lam) > SMEM_LIMIT) return fail(ctx, MRBF_EUNSUPPORTED, "population too large for the ranking kernel%s");
    CK(cudaSetDevice(ctx->device));
    const size_t pop = (size_t)B * lam * n;
    ENSURE(ctx->ws[15], sizeof(double) * (4 * pop + (size_t)B * lam * k) + sizeof(int) * (size_t)B * lam);
    PsParams P{};
    P.B = B; P.n = n; P.k = k; P.lambda = lam; P.mu = (lam + 6) / 7; P.generations = gens; P.n_obj = n_obj; P.objective = objective;
    P.seed = (unsigned long long)seed; P.x0 = x; P.lb = lb; P.ub = ub; P.mx = mx; P.dir = dir;
    double* base = (double*)ctx->ws[15].p;
    P.X = base; P.S = base + pop; P.Xn = base + 2 * pop; P.Sn = base + 3 * pop;
    double* Y = base + 4 * pop; P.Y = Y; P.rank = (int*)(Y + (size_t)B * lam * k);
    P.best_f = f_min; P.best_x = x_min; P.best_y = y_min; P.best_found = found;
    CK(launch_ps_init(P, ctx->stream));
    ctx->launches += 1;
    for (int g = 0; g <= gens; ++g) {
        int rc = mrbf_eval_dev(ctx, m, lam, P.X, Y, nullptr);
        if (rc != MRBF_OK) return rc;
        CK(launch_ps_fitness_rank(P, g, ctx->stream));
        ctx->launches += 1;
        if (g == gens) break;
        CK(launch_ps_evolve(P, g + 1, ctx->stream));
        ctx->launches += 1;
        double* t = P.X; P.X = P.Xn; P.Xn = t; t = P.S; P.S = P.Sn; P.Sn = t;
    }
    if (evals_used) *evals_used = (gens + 1) * lam;
    return MRBF_OK;
}

int mrbf_ps_solve(mrbf_ctx* ctx, const mrbf_model* m, const double* x, const double* lb, const double* ub,
                  const double* mx, const double* dir, int32_t n_obj, int32_t objective, int32_t population,
                  int32_t max_evals, int64_t seed, double* f_min, double* x_min, double* y_min, int32_t* found,
                  int32_t* evals_used) {
    if (!ctx || !m || !x || !lb || !ub || !f_min || !x_min || !y_min || !found) return MRBF_EINVAL;
    CK(cudaSetDevice(ctx->device));
    const int B = m->B, n = m->n, k = m->k;
    const size_t bx = sizeof(double) * (size_t)B * n, bk = sizeof(double) * (size_t)B * k, bb = sizeof(double) * (size_t)B;
    ENSURE(ctx->hb[23], 4 * bx + 3 * bk + bb + sizeof(int) * (size_t)B);
    double* d_x = (double*)ctx->hb[23].p; double* d_lb = d_x + (size_t)B * n; double* d_ub = d_lb + (size_t)B * n; double* d_xm = d_ub + (size_t)B * n;
    double* d_mx = d_xm + (size_t)B * n; double* d_dir = d_mx + (size_t)B * k; double* d_ym = d_dir + (size_t)B * k;
    double* d_f = d_ym + (size_t)B * k; int* d_found = (int*)(d_f + B);
    H2D(d_x, x, bx); H2D(d_lb, lb, bx); H2D(d_ub, ub, bx);
    if (mx) H2D(d_mx, mx, bk);
    if (dir) H2D(d_dir, dir, bk);
    int rc = mrbf_ps_solve_dev(ctx, m, d_x, d_lb, d_ub, mx ? d_mx : nullptr, dir ? d_dir : nullptr, n_obj, objective, population,
                               max_evals, seed, d_f, d_xm, d_ym, d_found, evals_used);
    if (rc != MRBF_OK) return rc;
    D2H(f_min, d_f, bb); D2H(x_min, d_xm, bx); D2H(y_min, d_ym, bk); D2H(found, d_found, sizeof(int) * (size_t)B);
    CK(cudaStreamSynchronize(ctx->stream));
    return MRBF_OK;
}

// ------------------------------------------------------------------------------------------------ device-resident database
int mrbf_db_append_dev(mrbf_ctx* ctx, int32_t B, int32_t n, int32_t k, int32_t db_stride, double* sites, double* values,
                       int32_t* n_db, int32_t add_stride, const double* new_sites, const double* new_values,
                       const int32_t* n_add, int32_t* first_id, int32_t* status) {
    if (!ctx || !sites || !values || !n_db || !new_sites || !n_add || !first_id) return MRBF_EINVAL;
    if (B <= 0 || n <= 0 || k <= 0 || db_stride <= 0 || add_stride <= 0) return fail(ctx, MRBF_EINVAL, "bad sizes%s");
    CK(cudaSetDevice(ctx->device));
    DbAppendParams P{};
    P.B = B; P.n = n; P.k = k; P.db_stride = db_stride; P.add_stride = add_stride;
    P.sites = sites; P.values = values; P.n_db = n_db; P.new_sites = new_sites; P.new_values = new_values; P.n_add = n_add;
    P.first_id = first_id; P.status = status;
    CK(launch_db_append(P, ctx->stream));
    ctx->launches += 1;
    return MRBF_OK;
}

int mrbf_model_scatter_dev(mrbf_ctx* ctx, mrbf_model* dst, const mrbf_model* src, const int32_t* map, int32_t S) {
    if (!ctx || !dst || !src || !map) return MRBF_EINVAL;
    if (S <= 0 || S > src->B) return fail(ctx, MRBF_EINVAL, "mrbf_model_scatter: S must be in 1..B(src)%s");
    if (dst->n != src->n || dst->k != src->k || dst->train_stride < src->train_stride || dst->p != src->p || dst->deg != src->deg ||
        dst->kernel != src->kernel || dst->ibeta != src->ibeta || dst->sgn != src->sgn)
        return fail(ctx, MRBF_EINVAL, "mrbf_model_scatter: the two model batches differ in shape or radial function%s");
    CK(cudaSetDevice(ctx->device));
    ModelScatterParams P{};
    P.S = S; P.B_dst = dst->B; P.n = src->n; P.k = src->k; P.train_stride = src->train_stride; P.dst_stride = dst->train_stride; P.pl = src->p > 0 ? src->p : 1;
    P.map = map; P.src_N = src->N; P.src_centers = src->centers; P.src_w = src->w; P.src_lam = src->lam; P.src_alpha2 = src->alpha2;
    P.dst_N = dst->N; P.dst_centers = dst->centers; P.dst_w = dst->w; P.dst_lam = dst->lam; P.dst_alpha2 = dst->alpha2;
    CK(launch_model_scatter(P, ctx->stream));
    dst->pack_valid = false;
    ctx->launches += 1;
    return MRBF_OK;
}

}  // extern "C"
