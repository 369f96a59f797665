// Parameter blocks and launcher prototypes shared by the kernel translation units and the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "mrbf_common.cuh"

namespace mrbf {

struct CfgDev {
    int polynomial_degree;
    int optimized_sampling;
    double theta_enlarge_1, theta_enlarge_2, theta_pivot;
};

struct SelectParams {
    int B, n, db_stride, found_stride;
    CfgDev cfg;
    double delta_max;
    double approx_rtol;           // rtol of isapprox(delta, delta_max), RbfModel.jl:588 (mrbf_set_isapprox_rtol)
    const double* sites; const int* n_db; const int* x_index; const double* x; const double* delta;
    const double* glb; const double* gub; const int* flags_in; const int* max_new;
    int* r1; int* n_r1; int* r2; int* n_r2; double* r3_sites; int* n_r3; double* dirs; int* n_dirs; int* flags_out;
    // workspace / hand-over to round 4
    double* S; double* T; unsigned char* cflags; double* WZ; int wz_in_smem; int st_in_smem;
    double* lb2; double* ub2; int* found; int* n_found;
    long long* dbg_clock;         // instrumentation (MRBF_DEBUG_CLOCK): phase time stamps of CTA 0, or NULL
};

struct Round4Params {
    int B, n, db_stride, found_stride, extra_stride, r4_stride, NM, max_points;
    CfgDev cfg;
    RadFn rf;
    double chol_thr;
    const double* sites; const int* n_db; const double* lb2; const double* ub2;
    const int* found; const int* n_found; const double* extra_sites; const int* n_extra;
    int* r4; int* n_r4; int* status;
    unsigned char* cand;          // B x db_stride candidate flags (workspace)
    const unsigned char* cflags;  // flag bytes of select_rounds123_kernel (bit 2: in box 2, bit 4: picked) or NULL
    double* ws; size_t ws_stride; int ws_in_smem;
    int b0;                       // first instance of this launch (chunked literal launches)
    int only_marked;              // literal kernel: process only instances the fast kernel marked (n_r4 == -1)
    double* fs; size_t fs_stride; int fs_in_smem;   // fast-path state
    double* keep_fs; int* elig;                     // kept factorisation (mrbf_prepared) or NULL
    double* panel_ws;                               // round4_prep_kernel -> round4_elim_kernel hand-over (B x SchurGeom::pw_doubles)
    long long* dbg_clock;                           // instrumentation (MRBF_DEBUG_CLOCK): phase time stamps of CTA 0, or NULL
    // build mode of the register-tiled kernels (mrbf_build without a kept factorisation): `sites` are the training sets
    // (B x db_stride x n, n_db = N), the found set is the first p training points, every other point is a candidate that MUST be
    // accepted (plain Cholesky of the reduced kernel matrix, no leverages); instances that do not qualify get n_r4 = -1
    int build_mode;
    const double* shape_arr; double alpha_default; double* alpha2_out;   // per-instance shape parameter (NaN: default) or NULL
    int* found_out; int* n_found_out;               // build mode: the found ids 1..p are written here for build_schur_kernel    // Hand-over of under-poised instances (N0 < p: budget-limited round 3).  The literal kernel, launched first in `prefix_mode`, walks
    // such an instance only until the point set is poised (N = p: the rank guard of RbfModel.jl:433-438 no longer applies), leaves its
    // acceptances in r4[0 .. pre_cnt) and the first untried id in pre_min; the register kernels then continue with found set =
    // found | extra | r4[0 .. pre_cnt) and the candidates >= pre_min.  Decision-equivalent to the reference's walk only for kernels of
    // cpd order <= 1 (run_round4 in mrbf_api.cu explains why and enables it only for those).
    // hyb[b]: 0 regular instance, 1 hand-over pending, 2 finished by the literal kernel (never became poised).  All NULL: old behaviour.
    int prefix_mode;
    int* hyb; int* pre_cnt; int* pre_min;
};

// Geometry of the register-tiled round-4 kernel (mrbf_round4_schur.cu): shared-memory offsets and the layout of the
// kept factorisation, all in doubles.
struct SchurGeom {
    int MC, TR, ntiles, nthreads, eligible, two_variants, LD_small;
    size_t smem_doubles, smem_small;                                          // kernel 2 (panels + tiles + elimination): full / small launch shape
    size_t ps_Aq, ps_X0, ps_red, ps_int, ps_doubles;                          // kernel 1 (candidate list, Pi_0^{-1})
    size_t pw_M0, pw_clist, pw_meta, pw_doubles;                              // global hand-over workspace per instance
    size_t off_M0, off_U, off_C, off_L, off_acc, state_doubles;
};

struct GatherParams {
    int B, n, k, db_stride, r4_stride, train_stride;
    const double* sites; const double* values; const int* x_index;
    const int* r1; const int* n_r1; const int* r2; const int* n_r2;
    const double* r3_sites; const double* r3_values; const int* n_r3; const int* r4; const int* n_r4;
    double* train_sites; double* train_values; int* N;
    const int* skip;      // instances to leave untouched (or NULL)
};

struct BuildParams {
    int B, n, k, train_stride, p, deg;
    int kernel, ibeta; double sgn;        // RadFn pieces; alpha2 is per instance
    double alpha_default;
    const int* N; const double* sites; const double* values; const double* shape;
    double* w; double* lam; double* alpha2_out; int* status;
    double* ws; size_t ws_stride; int ws_in_smem; int ld; int smem_ws_doubles;
    int stage_off;        // > 0: offset (doubles) of a coordinate-major staging area for the sites behind the in-smem system
    const int* skip;      // instances already built by the prepared path (or NULL)
    double* centers_out;  // when non-NULL the kernel also copies the training sites into the model
    int* N_out;
};

struct PreparedBuildParams {
    int B, n, k, NM, p, deg, db_stride, found_stride, r4_stride, train_stride;
    size_t fs_stride, off_M0, off_G, off_C, off_L;
    const double* fs; const int* elig; const int* found; const int* n_found; const int* n_extra; const int* r4; const int* n_r4;
    const double* values; const double* r3_values;
    double alpha2;
    double* centers; double* w; double* lam; double* alpha2_out; int* N; int* status; int* done;
};

// Build from the factorisation kept by round4_elim_kernel (layout: SchurGeom::off_*).
struct SchurBuildParams {
    int B, n, k, p, deg, db_stride, found_stride, r4_stride, train_stride, MC;
    size_t fs_stride, off_M0, off_U, off_C, off_L, off_acc;
    const double* fs; const int* elig; const int* found; const int* n_found; const int* r4;
    const double* sites; const double* values; const double* r3_sites; const double* r3_values;
    double alpha2;        // < 0: alpha2_out has already been written per instance (build mode)
    double* centers; double* w; double* lam; double* alpha2_out; int* N; int* status; int* done;
};

struct EvalParams {
    int B, n, k, train_stride, p, deg;
    int kernel, ibeta; double sgn;
    long long M;
    const int* N; const double* centers; const double* w; const double* lam; const double* alpha2;
    const double* X; double* Y; double* J;
    // tiled, centred copy of the model for the DMMA kernel (built once per model): per instance nt tiles of
    // [64 x s centre rows | 64 squared norms | k x 64 coefficients]
    const double* pack; int pack_s, pack_nt; size_t pack_tile_doubles;
    // small grids (few trial points x few instances): the centre tiles are split over gridDim.z CTAs per point tile, partial sums
    // go to partY / partJ (zsplit x the output shapes) and eval_split_reduce_kernel adds them in a fixed order plus the tail
    int zsplit; double* partY; double* partJ;
};

struct PackParams {
    int B, n, k, train_stride, s, nt; size_t tile_doubles;
    const int* N; const double* centers; const double* w; double* pack;
};

struct BacktrackParams {
    int B, n, k, nsteps, strict;
    double armijo_c, shrink, min_stepsize; int max_loops;
    const double* x; const double* dir; const double* step0; const double* omega;
    const double* Yall;                    // B x (nsteps + 1) x k : m(x) then m(x + sigma_i dir)
    double* Xall;                          // B x (nsteps + 1) x n
    double* sig_all;                       // B x nsteps
    int* step_index; double* sigma; double* x_plus; double* mx; double* mx_plus;
};

struct DescentParams {
    int B, n, k, normalize; size_t warp_doubles;
    const double* jac; const double* x; const double* lb; const double* ub;
    double* d; double* omega; int* iters; int* status;
};

// Pascoletti-Serafini / ideal-point inner solves (mrbf_ps.cu)
struct PsParams {
    int B, n, k, lambda, mu, generations, n_obj, objective;
    unsigned long long seed;
    const double* x0; const double* lb; const double* ub; const double* mx; const double* dir;
    double* X; double* S; double* Xn; double* Sn;       // populations and step sizes, B x lambda x n each
    const double* Y;                                    // surrogate values of X, B x lambda x k
    int* rank;                                          // B x lambda
    double* best_f; double* best_x; double* best_y; int* best_found;
};

struct DbAppendParams {
    int B, n, k, db_stride, add_stride;
    double* sites; double* values; int* n_db;
    const double* new_sites; const double* new_values; const int* n_add;
    int* first_id; int* status;
};

struct ModelScatterParams {
    int S, B_dst, n, k, train_stride, dst_stride, pl;
    const int* map;
    const int* src_N; const double* src_centers; const double* src_w; const double* src_lam; const double* src_alpha2;
    int* dst_N; double* dst_centers; double* dst_w; double* dst_lam; double* dst_alpha2;
};

size_t select_smem_bytes(int n, bool wz_in_smem, int st_doubles, int db_stride);
// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the FUNCTION, process-wide: two host threads that launch the same kernel with
// different shared-memory sizes through their own contexts (the reference's benchmark driver runs optimize() under Threads.@threads,
// examples/large_scale_benchmarks.jl:253) must not lower each other's limit between the set and the launch.  The limit is therefore
// only ever raised, under a mutex (mrbf_api.cu); sizes within the 48 KB default need no call at all.
cudaError_t raise_dyn_smem_impl(const void* kernel, size_t bytes);
template <class... A> inline cudaError_t raise_dyn_smem(void (*kernel)(A...), size_t bytes) { return raise_dyn_smem_impl((const void*)kernel, bytes); }

size_t round4_vec_doubles(int n, int NM, int p);
size_t round4_ws_doubles(int n, int NM, int p);
size_t round4_fast_state_doubles(int n, int NM, int p);
void round4_fast_state_layout(int n, int NM, int p, size_t* off_M0, size_t* off_G, size_t* off_C, size_t* off_L);
size_t build_vec_doubles(int n, int k, int ld, int p);
size_t build_ws_doubles(int n, int k, int ld, int p);

cudaError_t launch_select_rounds123(const SelectParams& P, size_t smem, cudaStream_t s);
bool select_mma_eligible(int n, int db_stride);                       // register/DMMA filter kernel (mrbf_select_mma.cu)
cudaError_t launch_select_rounds123_mma(const SelectParams& P, cudaStream_t s);
cudaError_t launch_round4(const Round4Params& P, size_t smem, cudaStream_t s, int grid);
cudaError_t launch_round4_block(const Round4Params& P, int T, size_t smem, cudaStream_t s);
size_t round4_block_vec_doubles(int T, int n, int NM, int p);
SchurGeom round4_schur_geom(int n, int p, int db_stride);
cudaError_t launch_round4_schur(const Round4Params& P, const SchurGeom& g, cudaStream_t s);
cudaError_t launch_gather_training(const GatherParams& P, cudaStream_t s);
cudaError_t launch_build(const BuildParams& P, size_t smem, cudaStream_t s);
cudaError_t launch_build_prepared(const PreparedBuildParams& P, size_t smem, cudaStream_t s);
size_t build_prepared_smem_doubles(int n, int k, int NM, int p);
cudaError_t launch_build_prepared_stream(const PreparedBuildParams& P, size_t smem, cudaStream_t s);
size_t build_prepared_stream_smem_doubles(int n, int k, int NM, int p);
cudaError_t launch_build_schur(const SchurBuildParams& P, size_t smem, cudaStream_t s);
size_t build_schur_smem_doubles(int k, int MC, int p);
cudaError_t launch_eval(const EvalParams& P, cudaStream_t s, int* n_launches);
cudaError_t launch_eval_pack(const PackParams& P, cudaStream_t s);
int eval_pack_stride(int n);
int eval_split_factor(long long M, int B, int pack_nt);
cudaError_t launch_descent_direction(const DescentParams& P, cudaStream_t s);
size_t descent_warp_doubles(int n, int k);
int descent_max_outputs();
cudaError_t launch_backtrack_points(const BacktrackParams& P, cudaStream_t s);
cudaError_t launch_backtrack_pick(const BacktrackParams& P, cudaStream_t s);
cudaError_t launch_ps_init(const PsParams& P, cudaStream_t s);
cudaError_t launch_ps_fitness_rank(const PsParams& P, int gen, cudaStream_t s);
cudaError_t launch_ps_evolve(const PsParams& P, int gen, cudaStream_t s);
size_t ps_rank_smem_bytes(int lambda);
cudaError_t launch_db_append(const DbAppendParams& P, cudaStream_t s);
cudaError_t launch_model_scatter(const ModelScatterParams& P, cudaStream_t s);

}  // namespace mrbf
