// Pascoletti-Serafini inner solves on the surrogates (src/descent.jl:369-387 `_min_component`, 404-412 `compute_local_ideal_point`,
// 478-581 `_ps_optimization` / `get_criticality(::PascolettiSerafiniConfig)`), batched over instances and over the population.
//
// The reference hands both problems to NLopt (un-vendored dependency NLopt.jl / nlopt 2.7, `:GN_ISRES`, `maxeval = 500 (n + 1)`,
// `xtol_rel = 1e-3`) and NLopt calls the surrogate one point at a time (AbstractSurrogateInterface.jl:98-106): 10^4-10^5 sequential
// evaluations per iteration.  ISRES is Runarsson & Yao's (mu, lambda) evolution strategy with stochastic ranking; it is restated here
// (published algorithm, not NLopt's source) so that a whole GENERATION of every instance is ONE batched surrogate evaluation:
//
//   generation g:   ps_fitness_rank_kernel   fitness + constraint penalty of the lambda individuals, stochastic ranking, best-so-far
//                   ps_evolve_kernel         lambda offspring of the mu = ceil(lambda / 7) best: differential variation for the
//                                            first mu - 1, log-normal self-adaptive mutation (with smoothing) for the rest
//                   mrbf_eval (values)       B x lambda trial points through the evaluation kernels (mrbf_eval.cu)
//
// Problems (mode):
//   ideal point   min  m_l(xi)                              s.t. lb <= xi <= ub, c(xi) <= 0        (descent.jl:369-387)
//   PS            min  tau   s.t.  m_l(xi) - m_l(x) - tau r_l <= 0 (l < n_obj),  -1 <= tau <= 0,  lb <= xi <= ub,  c(xi) <= 0
// For PS the variable tau is eliminated analytically: for a given xi the best feasible tau is
//   tau(xi) = clamp(max_l (m_l(xi) - m_l(x)) / r_l, -1, 0),
// so an individual is xi alone, its fitness is tau(xi) and its penalty the squared violation that is left (only when the clamp at 0
// binds, i.e. xi is worse than x) plus the constraint surrogates' -- same optimum, one search dimension less than NLopt sees.
// Random numbers are counter-based (splitmix64 of (seed, instance, generation, individual, coordinate, stream)), so a run is
// reproducible and independent of the launch geometry; parity with NLopt's own RNG stream is not attainable --
// the bar is solution quality against a CPU restatement and a multi-start SQP reference (tests/test_gpu_ps.py).
#include "mrbf_common.cuh"
#include "mrbf_kernels.h"

namespace mrbf {

namespace {

constexpr double PS_ALPHA = 0.2;      // smoothing of the step sizes
constexpr double PS_GAMMA = 0.85;     // differential variation
constexpr double PS_PF = 0.45;        // probability of comparing by fitness although a partner is infeasible
constexpr int PS_RETRY = 10;          // resampling of a mutation that leaves the box

__host__ __device__ inline unsigned long long splitmix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// key of one draw: (seed, instance, generation, individual, coordinate, stream)
__device__ __forceinline__ unsigned long long ps_key(unsigned long long seed, int b, int gen, int i, int j, int stream) {
    unsigned long long h = splitmix64(seed ^ (0xD1B54A32D192ED03ull * (unsigned long long)(b + 1)));
    h = splitmix64(h ^ ((unsigned long long)(unsigned)gen << 32 | (unsigned)i));
    return splitmix64(h ^ ((unsigned long long)(unsigned)j << 8 | (unsigned)stream));
}
__device__ __forceinline__ double u01(unsigned long long h) { return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }   // (0, 1)
__device__ __forceinline__ double normal01(unsigned long long h) {                                                                  // Box-Muller
    const double u1 = u01(h), u2 = u01(splitmix64(h));
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

}  // namespace

// Generation 0: individual 0 is the start point, the others are uniform in the box; step sizes (ub - lb) / sqrt(n).
__global__ void __launch_bounds__(256) ps_init_kernel(PsParams P) {
    const int b = blockIdx.x, n = P.n, lam = P.lambda;
    const double* lb = P.lb + (size_t)b * n; const double* ub = P.ub + (size_t)b * n; const double* x0 = P.x0 + (size_t)b * n;
    double* X = P.X + (size_t)b * lam * n; double* S = P.S + (size_t)b * lam * n;
    const double rs = rsqrt((double)n);
    for (int e = threadIdx.x; e < lam * n; e += blockDim.x) {
        const int i = e / n, j = e % n;
        const double w = ub[j] - lb[j];
        double v = (i == 0) ? x0[j] : lb[j] + u01(ps_key(P.seed, b, 0, i, j, 0)) * w;
        v = fmin(fmax(v, lb[j]), ub[j]);
        X[e] = v; S[e] = w * rs;
    }
    if (threadIdx.x == 0) { P.best_f[b] = INFINITY; P.best_found[b] = 0; }
}

// Fitness and penalty of the population, stochastic ranking (Runarsson & Yao) as an odd-even transposition sort -- lambda phases,
// each comparing disjoint neighbour pairs in parallel: by fitness when both are feasible or with probability PF, else by penalty --
// and the best feasible individual seen so far.
__global__ void __launch_bounds__(256) ps_fitness_rank_kernel(PsParams P, int gen) {
    extern __shared__ double smem[];
    const int b = blockIdx.x, n = P.n, k = P.k, lam = P.lambda, tid = threadIdx.x, nt = blockDim.x;
    double* f = smem; double* phi = f + lam; int* idx = reinterpret_cast<int*>(phi + lam);
    double* red = reinterpret_cast<double*>(idx + lam + (lam & 1));
    int* redi = reinterpret_cast<int*>(red + 40);
    const double* Y = P.Y + (size_t)b * lam * k;
    const double* mx = P.mx ? P.mx + (size_t)b * k : nullptr;
    const double* dir = P.dir ? P.dir + (size_t)b * k : nullptr;
    for (int i = tid; i < lam; i += nt) {
        const double* y = Y + (size_t)i * k;
        double fi, pen = 0.0;
        if (dir) {                                   // Pascoletti-Serafini: tau eliminated
            double t = -INFINITY;
            for (int l = 0; l < P.n_obj; ++l) t = fmax(t, (y[l] - mx[l]) / dir[l]);
            // tau = clamp(t, -1, 0).  t <= 0: every constraint holds at tau by construction (no residual is computed -- with a fused
            // multiply-add (y - mx) - (t r) would return the rounding error of the division and flag half the population).
            // t > 0: the point is worse than x in some objective; what is left at tau = 0 is the violation.  The start point itself
            // (individual 0 of generation 0) is feasible with tau = 0 whatever the rounding of its own evaluation says.
            if (t > 0.0 && !(gen == 0 && i == 0))
                for (int l = 0; l < P.n_obj; ++l) { const double v = y[l] - mx[l]; if (v > 0.0) pen = fma(v, v, pen); }
            fi = fmin(fmax(t, -1.0), 0.0);
        } else {
            fi = y[P.objective];
        }
        for (int l = P.n_obj; l < k; ++l) { const double v = y[l]; if (v > 0.0) pen = fma(v, v, pen); }      // constraint surrogates c <= 0
        if (!(fi == fi) || !(pen == pen)) { fi = INFINITY; pen = INFINITY; }
        f[i] = fi; phi[i] = pen; idx[i] = i;
    }
    __syncthreads();
    // best feasible individual of this generation (smallest fitness, smallest index on ties)
    {
        ArgMax m; m.v = 0.0; m.id = -1;              // reuse: maximise -f
        for (int i = tid; i < lam; i += nt) if (phi[i] == 0.0 && f[i] < INFINITY) { ArgMax c_; c_.v = -f[i]; c_.id = i; m = better(m, c_); }
        m = block_argmax(m, red, redi);
        if (m.id >= 0 && (-m.v < P.best_f[b] || !P.best_found[b])) {
            const double* xb = P.X + ((size_t)b * lam + m.id) * n;
            for (int j = tid; j < n; j += nt) P.best_x[(size_t)b * n + j] = xb[j];
            for (int l = tid; l < k; l += nt) P.best_y[(size_t)b * k + l] = Y[(size_t)m.id * k + l];
            __syncthreads();
            if (tid == 0) { P.best_f[b] = -m.v; P.best_found[b] = 1; }
        }
        __syncthreads();
    }
    if (gen == P.generations) return;                // the last evaluation only updates the best
    for (int ph = 0; ph < lam; ++ph) {
        for (int a = (ph & 1) + 2 * tid; a + 1 < lam; a += 2 * nt) {
            const int ia = idx[a], ib = idx[a + 1];
            const double u = u01(ps_key(P.seed, b, gen, ph, a, 1));
            bool swap;
            if ((phi[ia] == 0.0 && phi[ib] == 0.0) || u < PS_PF) swap = f[ia] > f[ib];
            else swap = phi[ia] > phi[ib];
            if (swap) { idx[a] = ib; idx[a + 1] = ia; }
        }
        __syncthreads();
    }
    int* rank = P.rank + (size_t)b * lam;
    for (int i = tid; i < lam; i += nt) rank[i] = idx[i];
}

// Offspring of the mu best (rank order).  Offspring i has parent rank[i mod mu].  i < mu - 1: differential variation
// x' = x_p + gamma (x_best - x_rank[i+1]) with the parent's step sizes; otherwise sigma' = sigma_p exp(tau' N_i + tau N_ij) (capped at
// (ub - lb) / sqrt(n)), x' = x_p + sigma' N(0, 1) resampled up to ten times while it leaves the box (else the parent's coordinate),
// and the step size is smoothed towards the parent's.  A differential step that leaves the box falls back to the mutation.
__global__ void __launch_bounds__(256) ps_evolve_kernel(PsParams P, int gen) {
    const int b = blockIdx.x, n = P.n, lam = P.lambda, mu = P.mu;
    const double* lb = P.lb + (size_t)b * n; const double* ub = P.ub + (size_t)b * n;
    const double* X = P.X + (size_t)b * lam * n; const double* S = P.S + (size_t)b * lam * n;
    double* Xn = P.Xn + (size_t)b * lam * n; double* Sn = P.Sn + (size_t)b * lam * n;
    const int* rank = P.rank + (size_t)b * lam;
    const double taup = 1.0 / sqrt(2.0 * (double)n), tau = 1.0 / sqrt(2.0 * sqrt((double)n)), rs = rsqrt((double)n);
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < lam * n; e += gridDim.y * blockDim.x) {
        const int i = e / n, j = e % n;
        const int par = rank[i % mu];
        const double xp = X[(size_t)par * n + j], sp = S[(size_t)par * n + j];
        const double lo = lb[j], hi = ub[j];
        double xv = xp, sv = sp;
        bool done = false;
        if (i < mu - 1) {
            const double cand = xp + PS_GAMMA * (X[(size_t)rank[0] * n + j] - X[(size_t)rank[i + 1] * n + j]);
            if (cand >= lo && cand <= hi) { xv = cand; done = true; }
        }
        if (!done) {
            const double gi = normal01(ps_key(P.seed, b, gen, i, n, 2));                 // one draw per individual
            double s1 = sp * exp(taup * gi + tau * normal01(ps_key(P.seed, b, gen, i, j, 3)));
            s1 = fmin(s1, (hi - lo) * rs);
            for (int t = 0; t < PS_RETRY; ++t) {
                const double cand = xp + s1 * normal01(ps_key(P.seed, b, gen, i, j, 4 + t));
                if (cand >= lo && cand <= hi) { xv = cand; break; }
            }
            sv = sp + PS_ALPHA * (s1 - sp);
        }
        Xn[e] = xv; Sn[e] = sv;
    }
}

size_t ps_rank_smem_bytes(int lambda) { return sizeof(double) * (2 * (size_t)lambda + 40 + 24) + sizeof(int) * ((size_t)lambda + (lambda & 1)); }

cudaError_t launch_ps_init(const PsParams& P, cudaStream_t s) {
    ps_init_kernel<<<P.B, 256, 0, s>>>(P);
    return cudaGetLastError();
}
cudaError_t launch_ps_fitness_rank(const PsParams& P, int gen, cudaStream_t s) {
    const size_t smem = ps_rank_smem_bytes(P.lambda);
    cudaError_t e = raise_dyn_smem(ps_fitness_rank_kernel, smem);
    if (e != cudaSuccess) return e;
    ps_fitness_rank_kernel<<<P.B, 256, smem, s>>>(P, gen);
    return cudaGetLastError();
}
cudaError_t launch_ps_evolve(const PsParams& P, int gen, cudaStream_t s) {
    int gy = (P.lambda * P.n + 256 * 8 - 1) / (256 * 8);
    if (gy < 1) gy = 1;
    if ((long long)P.B * gy < 296) gy = (296 + P.B - 1) / P.B;          // two CTAs per SM at least for small batches
    ps_evolve_kernel<<<dim3(P.B, gy), 256, 0, s>>>(P, gen);
    return cudaGetLastError();
}

}  // namespace mrbf
