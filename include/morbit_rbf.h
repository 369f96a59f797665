/* morbit_rbf.h -- C ABI of the B200-native RBF-surrogate hot path of Morbit.jl.
 *
 * This is the drop-in boundary: a Julia `GpuRbfConfig <: AbstractSurrogateConfig` (julia/GpuRbf.jl,
 * INTEGRATION.md) implements Morbit's surrogate plugin interface
 * (src/AbstractSurrogateInterface.jl:6-79) and reaches CUDA only through these `ccall`-able
 * entry points.  Plain pointers and sizes only; no torch / CUDA types in any signature.
 *
 * Conventions
 *   - all floating point is IEEE double; all ids are 1-based int32 (Morbit's Int ids, narrowed);
 *   - "sites" are array-of-structs like the reference's Vector{SVector{n}}: site i of instance b
 *     starts at sites[(b*stride + i) * n];  this is also Julia's column-major n x stride x B array;
 *   - every call is batched over B independent instances (multistart runs / objective groups);
 *     B = 1 is the single-`optimize` case;
 *   - entry points without suffix take HOST pointers (borrowed for the duration of the call,
 *     wrap in GC.@preserve) and return after the results are in host memory;
 *     `_dev` entry points take DEVICE pointers, enqueue on the context's stream and return
 *     without synchronising (mrbf_sync waits);
 *   - return value 0 = ok, < 0 = error (mrbf_last_error gives the text); never aborts the process;
 *     per-instance numerical failures are reported in `status[b]` and do not poison the batch;
 *   - re-entrant: no global mutable state; one context per (host thread, device).
 */
#ifndef MORBIT_RBF_H
#define MORBIT_RBF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRBF_ABI_VERSION 2

/* src/models/RbfModel.jl:48-54 (RbfKernels), same order */
enum mrbf_kernel {
    MRBF_CUBIC = 0,
    MRBF_INV_MULTIQUADRIC = 1,
    MRBF_MULTIQUADRIC = 2,
    MRBF_THIN_PLATE_SPLINE = 3,
    MRBF_GAUSSIAN = 4
};

enum mrbf_status {
    MRBF_OK = 0,
    MRBF_EINVAL = -1,        /* bad argument */
    MRBF_ECUDA = -2,         /* CUDA runtime error */
    MRBF_ENOMEM = -3,
    MRBF_EUNSUPPORTED = -4,  /* e.g. use_max_points (random sampling), a kernel that needs a polynomial tail of degree > 2 */
    MRBF_ENUMERIC = -5       /* at least one instance failed numerically; see status[] */
};

/* Mirrors RbfConfig, src/models/RbfModel.jl:66-112 (String shape parameters are evaluated on
 * the Julia side, RbfModel.jl:135-143, and arrive here as numbers via `shape`). */
typedef struct mrbf_cfg {
    int32_t kernel;                /* enum mrbf_kernel */
    int32_t polynomial_degree;     /* -1, 0, 1 */
    double shape_parameter;        /* NaN => package default (RbfModel.jl:673) */
    double theta_enlarge_1;
    double theta_enlarge_2;
    double theta_pivot;
    double theta_pivot_cholesky;
    int32_t max_model_points;      /* <= 0 => (n+1)(n+2)/2  (RbfModel.jl:356) */
    int32_t use_max_points;        /* must be 0: the reference draws rand() points (RbfModel.jl:409-414) */
    int32_t optimized_sampling;
    int32_t reserved;
} mrbf_cfg;

typedef struct mrbf_ctx mrbf_ctx;
typedef struct mrbf_model mrbf_model;   /* device-resident batch of fitted models (replaces RbfModel.model,
                                           src/models/RbfModel.jl:33-38) */

/* ---- context ------------------------------------------------------------------------------- */
int mrbf_abi_version(void);
int mrbf_init(int device, mrbf_ctx** ctx);                 /* creates a private non-blocking stream */
int mrbf_set_stream(mrbf_ctx* ctx, void* cuda_stream);     /* run on the caller's cudaStream_t instead */
int mrbf_get_stream(const mrbf_ctx* ctx, void** cuda_stream);  /* the cudaStream_t the *_dev entry points enqueue on: a caller whose device
                                                              buffers are produced on another stream orders the two with events */
int mrbf_sync(mrbf_ctx* ctx);
/* rtol of the reference's `Δ ≈ Δ_max` test that skips round 2 (src/models/RbfModel.jl:588).  Julia's isapprox uses
 * max(sqrt(eps(T))) over the two argument types, and delta_max(algo_config) of the DEFAULT config is a Float32 literal
 * (src/AbstractConfigInterface.jl:31), so a run with default settings compares with sqrt(eps(Float32)) = 3.4526698e-4, a run with an
 * AlgorithmConfig{Float64} with sqrt(eps(Float64)) = 1.49e-8 (the library default).  The Julia shim passes
 * Base.rtoldefault(typeof(Δ), typeof(delta_max(ac)), 0). */
int mrbf_set_isapprox_rtol(mrbf_ctx* ctx, double rtol);
void mrbf_destroy(mrbf_ctx* ctx);
const char* mrbf_last_error(const mrbf_ctx* ctx);
/* kernels launched through this context since creation (the bench's gpu_launches counter) */
int64_t mrbf_launch_count(const mrbf_ctx* ctx);
/* Instrumentation (not part of the reference interface): when enabled, every kernel launch is bracketed by CUDA
 * events on the context's stream.  mrbf_profile_read synchronises and returns the device time in ms of the LAST
 * launch of each kernel class: ms[0] rounds 1-3, ms[1] round 4, ms[2] training-set gather, ms[3] build,
 * ms[4] eval/Jacobian (all passes of the last call), ms[5] round-4 fallback kernel, ms[6] build from a kept
 * factorisation, ms[7] round-4 literal prefix run of under-poised instances (until they are poised). */
int mrbf_profile_enable(mrbf_ctx* ctx, int32_t on);
int mrbf_profile_read(mrbf_ctx* ctx, double* ms8);

/* ---- training-set search: replaces prepare_update_model rounds 1-4 --------------------------
 * src/models/RbfModel.jl:518-655 (_rbf_round1 :242, _rbf_round2 :251, _rbf_round3 :269, _rbf_round4 :352),
 * src/models/AffinelyIndependentPoints.jl:4-106, src/Databases.jl:324-327, src/utilities.jl:126-221, 437-448.
 *
 *   sites      B x db_stride x n   database sites of each instance (scaled space), ids 1..n_db[b]
 *   n_db       B
 *   x_index    B                   id of the current iterate in its database
 *   x          B x n               current scaled iterate
 *   delta      B                   trust-region radius;  delta_max: algorithm's maximum radius
 *   glb, gub   n                   global scaled bounds (+-INFINITY when unbounded)
 *   flags_in   B x 2               [ensure_fully_linear, force_rebuild]
 *   max_new    B                   evaluation budget for round 3 (RbfModel.jl:613-618, computed by the host)
 * outputs
 *   r1, r2     B x n ids           round-1/2 picks in selection order;  n_r1, n_r2: B
 *   r3_sites   B x n x n           new sites of round 3 (host appends them: new_result!, ids n_db+1..); n_r3: B
 *   r4         B x r4_stride ids   round-4 picks in acceptance order;  n_r4: B
 *   dirs       B x n x n           improving directions (column c at dirs[(b*n + c)*n]); n_dirs: B
 *   flags_out  B x 2               [fully_linear, rebuilt (round 3 fell back to the coordinate rebuild, :634)]
 *   status     B                   0 ok
 */
int mrbf_select_points(mrbf_ctx* ctx, const mrbf_cfg* cfg, int32_t B, int32_t n, int32_t db_stride,
                       const double* sites, const int32_t* n_db, const int32_t* x_index, const double* x,
                       const double* delta, double delta_max, const double* glb, const double* gub,
                       const int32_t* flags_in, const int32_t* max_new,
                       int32_t* r1, int32_t* n_r1, int32_t* r2, int32_t* n_r2, double* r3_sites, int32_t* n_r3,
                       int32_t r4_stride, int32_t* r4, int32_t* n_r4, double* dirs, int32_t* n_dirs,
                       int32_t* flags_out, int32_t* status);
int mrbf_select_points_dev(mrbf_ctx* ctx, const mrbf_cfg* cfg, int32_t B, int32_t n, int32_t db_stride,
                           const double* sites, const int32_t* n_db, const int32_t* x_index, const double* x,
                           const double* delta, double delta_max, const double* glb, const double* gub,
                           const int32_t* flags_in, const int32_t* max_new,
                           int32_t* r1, int32_t* n_r1, int32_t* r2, int32_t* n_r2, double* r3_sites, int32_t* n_r3,
                           int32_t r4_stride, int32_t* r4, int32_t* n_r4, double* dirs, int32_t* n_dirs,
                           int32_t* flags_out, int32_t* status);

/* Same as mrbf_select_points_dev, and additionally keeps the round-4 factorisation of every instance on the device
 * (inverse Cholesky factor of the reduced kernel matrix, RbfModel.jl:394-396, 469-477) in an opaque handle.  The
 * reference discards these matrices and notes that keeping them would save work (RbfModel.jl:657-660);
 * mrbf_build_prepared_dev turns them into the model with two triangular mat-vecs per output.
 * *prepared must be NULL or a handle from an earlier call (reused when the shapes match, else freed and replaced).
 * With cfg->optimized_sampling == 0 there is no round 4 and nothing to keep (RbfModel.jl:564-569, 647-652): the handle then
 * only records the found set, and mrbf_build_prepared* builds [centre; r1; r2; r3] by the general route. */
typedef struct mrbf_prepared mrbf_prepared;
int mrbf_select_points_keep_dev(mrbf_ctx* ctx, const mrbf_cfg* cfg, int32_t B, int32_t n, int32_t db_stride,
                                const double* sites, const int32_t* n_db, const int32_t* x_index, const double* x,
                                const double* delta, double delta_max, const double* glb, const double* gub,
                                const int32_t* flags_in, const int32_t* max_new,
                                int32_t* r1, int32_t* n_r1, int32_t* r2, int32_t* n_r2, double* r3_sites, int32_t* n_r3,
                                int32_t r4_stride, int32_t* r4, int32_t* n_r4, double* dirs, int32_t* n_dirs,
                                int32_t* flags_out, int32_t* status, mrbf_prepared** prepared);
/* Host-pointer twins (what the Julia shim of a single optimize() run calls): same arguments, host memory. */
int mrbf_select_points_keep(mrbf_ctx* ctx, const mrbf_cfg* cfg, int32_t B, int32_t n, int32_t db_stride,
                            const double* sites, const int32_t* n_db, const int32_t* x_index, const double* x,
                            const double* delta, double delta_max, const double* glb, const double* gub,
                            const int32_t* flags_in, const int32_t* max_new,
                            int32_t* r1, int32_t* n_r1, int32_t* r2, int32_t* n_r2, double* r3_sites, int32_t* n_r3,
                            int32_t r4_stride, int32_t* r4, int32_t* n_r4, double* dirs, int32_t* n_dirs,
                            int32_t* flags_out, int32_t* status, mrbf_prepared** prepared);
void mrbf_free_prepared(mrbf_ctx* ctx, mrbf_prepared* prepared);
/* update_model from a kept factorisation.  values: B x db_stride x k (database values, same ids as `sites`),
 * r3_values: B x n x k values of the new round-3 sites (may be NULL when there are none).  Instances whose round 4
 * did not go through the shared-memory path (N0 != p, no round 4 at all) are built by the general route inside the
 * same call, so the result is always a complete model batch with training order [centre; r1; r2; r3; r4]. */
int mrbf_build_prepared_dev(mrbf_ctx* ctx, const mrbf_cfg* cfg, const mrbf_prepared* prepared, int32_t k,
                            const double* sites, const double* values, const double* r3_sites, const double* r3_values,
                            const int32_t* x_index, const int32_t* r1, const int32_t* n_r1, const int32_t* r2, const int32_t* n_r2,
                            const int32_t* n_r3, mrbf_model** model, int32_t* status);

int mrbf_build_prepared(mrbf_ctx* ctx, const mrbf_cfg* cfg, const mrbf_prepared* prepared, int32_t k,
                        const double* sites, const double* values, const double* r3_sites, const double* r3_values,
                        const int32_t* x_index, const int32_t* r1, const int32_t* n_r1, const int32_t* r2, const int32_t* n_r2,
                        const int32_t* n_r3, mrbf_model** model, int32_t* status);

/* _rbf_round4 alone with an explicit found set (used after _exploit_other_rbf_metas!, RbfModel.jl:311-342, 562,
 * and called directly by test/rbf_models.jl:74-86).  found: B x found_stride ids (centre first), n_found: B;
 * extra_sites: B x extra_stride x n sites that are in the found set but not (yet) in `sites` (may be NULL);
 * lb2, ub2: B x n box. */
int mrbf_round4(mrbf_ctx* ctx, const mrbf_cfg* cfg, int32_t B, int32_t n, int32_t db_stride,
                const double* sites, const int32_t* n_db, const double* lb2, const double* ub2,
                int32_t found_stride, const int32_t* found, const int32_t* n_found,
                int32_t extra_stride, const double* extra_sites, const int32_t* n_extra,
                int32_t r4_stride, int32_t* r4, int32_t* n_r4, int32_t* status);

/* Device-side gather of the training set [centre; r1; r2; r3; r4] (_collect_indices, RbfModel.jl:178-186,
 * and :754-757).  values: B x db_stride x k; r3_values: B x n x k (values of the new round-3 sites).
 * Outputs train_sites B x train_stride x n, train_values B x train_stride x k, N: B. */
int mrbf_gather_training_dev(mrbf_ctx* ctx, int32_t B, int32_t n, int32_t k, int32_t db_stride,
                             const double* sites, const double* values, const int32_t* x_index,
                             const int32_t* r1, const int32_t* n_r1, const int32_t* r2, const int32_t* n_r2,
                             const double* r3_sites, const double* r3_values, const int32_t* n_r3,
                             int32_t r4_stride, const int32_t* r4, const int32_t* n_r4,
                             int32_t train_stride, double* train_sites, double* train_values, int32_t* N);

/* ---- model build: replaces update_model -> RBF.RBFInterpolationModel ---------------------------
 * src/models/RbfModel.jl:743-767.  sites B x train_stride x n, values B x train_stride x k, N: B (sites used).
 * shape: B per-instance shape parameters or NULL (then cfg->shape_parameter).  Creates one device-resident
 * handle for the whole batch.  *model must be NULL or a handle from an earlier build call: its device buffers are
 * recycled when the shapes match (no allocation on the hot path; the reference likewise replaces the model in
 * place, SurrogateContainer.jl:376-382), else it is freed and replaced.  status[b]: 0 ok, > 0 reduced kernel matrix not
 * positive definite at that column.
 * Ownership rule of every mrbf_build* entry point: the handle passed in through *model belongs to the call.  MRBF_OK and
 * MRBF_ENUMERIC (host-pointer twins only: at least one instance failed numerically, the others are fine -- read status[])
 * return a VALID handle in *model; every other error releases the handle and leaves *model NULL. */
int mrbf_build(mrbf_ctx* ctx, const mrbf_cfg* cfg, int32_t B, int32_t n, int32_t k, int32_t train_stride,
               const int32_t* N, const double* sites, const double* values, const double* shape,
               mrbf_model** model, int32_t* status);
int mrbf_build_dev(mrbf_ctx* ctx, const mrbf_cfg* cfg, int32_t B, int32_t n, int32_t k, int32_t train_stride,
                   const int32_t* N, const double* sites, const double* values, const double* shape,
                   mrbf_model** model, int32_t* status);
void mrbf_free_model(mrbf_ctx* ctx, mrbf_model* model);
/* dimensions: out[0..5] = B, n, k, train_stride, p (polynomial basis size), effective polynomial degree */
int mrbf_model_dims(const mrbf_model* model, int32_t* out6);
/* coefficients for parity tests: w B x train_stride x k, lambda B x p x k (row c = basis function c: 1, x_1..x_n) */
int mrbf_model_coeffs(mrbf_ctx* ctx, const mrbf_model* model, double* w, double* lambda);

/* ---- evaluation: replaces eval_models / get_gradient / get_jacobian -----------------------------
 * src/models/RbfModel.jl:783-800.  X: B x M x n trial points (M per instance);
 * Y: B x M x k values; J: B x M x k x n Jacobians (row l = gradient of output l).  Y or J may be NULL. */
int mrbf_eval(mrbf_ctx* ctx, const mrbf_model* model, int64_t M, const double* X, double* Y, double* J);
int mrbf_eval_dev(mrbf_ctx* ctx, const mrbf_model* model, int64_t M, const double* X, double* Y, double* J);

/* Armijo backtracking over surrogate values, all step sizes in one launch (src/descent.jl:150-185):
 * evaluates m(x) and m(x + sigma_i * dir) for sigma_i = step0 * shrink^i (i = 0..max_loops, built by repeated
 * multiplication like the reference) and returns the first i that satisfies the (strict) Armijo condition
 * all(mx - mx_i >= sigma_i * c * omega), or the i at which the reference loop would have stopped.
 * x, dir: B x n;  step0, omega: B;  outputs: step_index B, sigma B, x_plus B x n, mx B x k, mx_plus B x k. */
int mrbf_backtrack(mrbf_ctx* ctx, const mrbf_model* model, const double* x, const double* dir, const double* step0,
                   const double* omega, double armijo_c, double shrink, double min_stepsize, int32_t max_loops,
                   int32_t strict, int32_t* step_index, double* sigma, double* x_plus, double* mx, double* mx_plus);
/* same with device pointers (all outputs required), enqueued on the context's stream */
int mrbf_backtrack_dev(mrbf_ctx* ctx, const mrbf_model* model, const double* x, const double* dir, const double* step0,
                       const double* omega, double armijo_c, double shrink, double min_stepsize, int32_t max_loops,
                       int32_t strict, int32_t* step_index, double* sigma, double* x_plus, double* mx, double* mx_plus);

/* ---- steepest-descent direction: replaces _steepest_descent_direction (src/descent.jl:75-135, JuMP + OSQP) -----------
 *      min alpha  s.t.  Df_i . d <= alpha * ||Df_i||_2 (rows normalised iff `normalize`),  -1 <= d <= 1,  lb <= x + d <= ub
 * solved exactly (bounded-variable simplex, k x k basis) for B instances at once.
 *   jac B x k x n (row l = gradient of output l, as mrbf_eval returns it), x B x n, lb / ub: n global scaled bounds
 *   (+-INFINITY when unbounded).  Outputs: d B x n, omega B (= -alpha, the criticality measure), iters B (may be NULL),
 *   status B (0 optimal, 1 iteration limit).  k <= 8.  Linear constraints of the MOP (A_eq, A_ineq) are not supported. */
int mrbf_descent_direction(mrbf_ctx* ctx, int32_t B, int32_t n, int32_t k, const double* jac, const double* x,
                           const double* lb, const double* ub, int32_t normalize,
                           double* d, double* omega, int32_t* iters, int32_t* status);
int mrbf_descent_direction_dev(mrbf_ctx* ctx, int32_t B, int32_t n, int32_t k, const double* jac, const double* x,
                               const double* lb, const double* ub, int32_t normalize,
                               double* d, double* omega, int32_t* iters, int32_t* status);

/* ---- Pascoletti-Serafini inner solves: replace the NLopt runs of src/descent.jl:369-387 (`_min_component`, called by
 * `compute_local_ideal_point` :404-412) and :478-500 (`_ps_optimization`, called by `get_criticality(::PascolettiSerafiniConfig)`
 * :503-581), whose objective / constraint callbacks evaluate the surrogates one point at a time
 * (src/AbstractSurrogateInterface.jl:98-106).  A batched (mu, lambda) evolution strategy with stochastic ranking (ISRES, the
 * reference's default `:GN_ISRES`): every generation of all B instances is one batched surrogate evaluation on the device.
 *   model: B instances, k outputs; outputs 0..n_obj-1 are objectives, n_obj..k-1 constraint surrogates c(xi) <= 0
 *          (`get_nl_ineq_constraints_optim_handles`).
 *   x B x n: start point (scaled iterate);  lb / ub B x n: `local_bounds(scal, x, delta)` (finite).
 *   dir == NULL (ideal-point mode): minimise output `objective` over the box (one call per objective, descent.jl:408-411);
 *          f_min[b] = minimum found.
 *   dir != NULL (PS mode): mx B x k = m(x), dir B x k = r > 0 (only the first n_obj entries are read); solves
 *          min tau  s.t.  m_l(xi) - mx_l - tau r_l <= 0,  -1 <= tau <= 0, box, c(xi) <= 0;  f_min[b] = tau (omega = |tau|).
 *   population <= 0: 20 (n + 1) (NLopt's ISRES default); max_evals <= 0: 500 (n + 1) (descent.jl:373, 418, 535).
 *   Outputs: x_min B x n, y_min B x k (surrogate values at x_min), found B (0: no feasible point was seen -- the caller
 *   treats the instance like NLopt's FAILURE, descent.jl:556-571), *evals_used (may be NULL): surrogate evaluations per instance.
 * Deterministic for a given seed (counter-based random numbers); NLopt's own random stream is not reproduced. */
int mrbf_ps_solve(mrbf_ctx* ctx, const mrbf_model* model, const double* x, const double* lb, const double* ub,
                  const double* mx, const double* dir, int32_t n_obj, int32_t objective, int32_t population,
                  int32_t max_evals, int64_t seed, double* f_min, double* x_min, double* y_min, int32_t* found,
                  int32_t* evals_used);
int mrbf_ps_solve_dev(mrbf_ctx* ctx, const mrbf_model* model, const double* x, const double* lb, const double* ub,
                      const double* mx, const double* dir, int32_t n_obj, int32_t objective, int32_t population,
                      int32_t max_evals, int64_t seed, double* f_min, double* x_min, double* y_min, int32_t* found,
                      int32_t* evals_used);

/* ---- device-resident database + model swap (lock-step multistart driver) ----------------------
 * new_result!(db, x, y) for B databases that live on the device (src/Databases.jl:174-183; value-less results are the
 * "unevaluated" ones of :202-205 and are stored as NaN rows): instance b appends its first n_add[b] rows of
 * new_sites (B x add_stride x n) / new_values (B x add_stride x k, or NULL => NaN) behind its n_db[b] sites, ids
 * n_db[b]+1 .. ; first_id[b] = id of the first appended row; n_db is updated in place.  status[b] = 1 (and nothing is
 * written, first_id[b] = 0) when the instance's capacity db_stride would be exceeded.  The box scan of
 * results_in_box_indices (:324-327) runs inside mrbf_select_points*_dev on the same buffers, so an iteration uploads
 * nothing but the values of the newly evaluated sites. */
int mrbf_db_append_dev(mrbf_ctx* ctx, int32_t B, int32_t n, int32_t k, int32_t db_stride, double* sites, double* values,
                       int32_t* n_db, int32_t add_stride, const double* new_sites, const double* new_values,
                       const int32_t* n_add, int32_t* first_id, int32_t* status);
/* Replace instances of a model batch by instances of another (the container swaps a rebuilt model in,
 * src/SurrogateContainer.jl:376-382): instance s < S of `src` becomes instance map[s] of `dst` (map[s] < 0: skipped).
 * Lets a lock-step driver rebuild a subset of the instances (criticality loop, model-improvement steps) in a compact
 * batch.  Both batches must have the same n, k, polynomial tail and radial function; `dst` may have a larger train_stride
 * (room for more training points per instance) than `src`. */
int mrbf_model_scatter_dev(mrbf_ctx* ctx, mrbf_model* dst, const mrbf_model* src, const int32_t* map, int32_t S);

/* ---- multi-GPU: the final gather of per-instance results (the only collective of the path) -----------------
 * Independent multistart instances / objective blocks shard over the GPUs of a box with NO collective on the data path (the
 * reference runs them independently under Threads.@threads, examples/large_scale_benchmarks.jl:253).  When the runs are over every
 * rank (one host process or thread per GPU) contributes its result rows -- x, f(x), stop code, #evals, training ids ... -- and
 * receives everybody's: one ncclAllGather over NVLink / NVSwitch.  NCCL is loaded at run time (dlopen "libnccl.so.2", or the path in
 * $MRBF_NCCL_LIB); without it these entry points return MRBF_EUNSUPPORTED and nothing else of the library is affected.
 *
 *   mrbf_comm_unique_id   rank 0 creates the 128-byte NCCL id; the host distributes it (Distributed.jl, MPI, a file ...)
 *   mrbf_comm_init        ncclCommInitRank on `device` (collective: every rank calls it with the same id)
 *   mrbf_comm_from_nccl   wrap a communicator the host already has (NCCL.jl, torch) -- not destroyed with the handle
 *   mrbf_gather           rows: count x width doubles of this rank (HOST pointer); all_rows: world x max_count x width doubles,
 *                         block r holds rank r's rows (zero padded), counts[r] its row count; max_count >= every rank's count and
 *                         equal on all ranks.  Blocks until the result is in host memory. */
typedef struct mrbf_comm mrbf_comm;
int mrbf_comm_unique_id(char* id128);
int mrbf_comm_init(int device, const char* id128, int32_t rank, int32_t world, mrbf_comm** comm);
int mrbf_comm_from_nccl(int device, void* nccl_comm, int32_t rank, int32_t world, mrbf_comm** comm);
void mrbf_comm_destroy(mrbf_comm* comm);
const char* mrbf_comm_last_error(const mrbf_comm* comm);
int mrbf_gather(mrbf_comm* comm, const double* rows, int32_t count, int32_t width, int32_t max_count, double* all_rows, int32_t* counts);

#ifdef __cplusplus
}
#endif
#endif /* MORBIT_RBF_H */
