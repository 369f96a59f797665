/* CPU oracle, C restatement of Morbit.jl's RBF-surrogate hot path (float64).
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py, never by the product library.
 *
 * PARITY UNPINNED (see oracle/rbf_oracle.py header): no Julia binary and no golden vectors
 * exist for this path; this file is pinned against oracle/rbf_oracle.py (NumPy + LAPACK,
 * literal dense restatement) by tests/test_oracle_c_vs_py.py and by the properties of
 * /root/reference/test/rbf_models.jl.
 *
 * Follows (reference file:line):
 *   results_in_box_indices            src/Databases.jl:324-327
 *   _orthogonal_complement_matrix     src/models/AffinelyIndependentPoints.jl:4-11
 *   AffinelyIndependentPointFilter    src/models/AffinelyIndependentPoints.jl:51-106
 *   _intersect_bounds (:absmax)       src/utilities.jl:126-221, 285-287
 *   nullify_last_row                  src/utilities.jl:437-448
 *   _rbf_round1/2/3, _rbf_round4      src/models/RbfModel.jl:205-307, 352-499
 *   prepare_update_model              src/models/RbfModel.jl:518-655
 *   update_model / RBFInterpolationModel   src/models/RbfModel.jl:743-767 (+ dependency, restated)
 *   eval_models/get_gradient/get_jacobian  src/models/RbfModel.jl:783-800 (+ dependency, restated)
 *
 * Difference in *cost* from the reference (results identical up to rounding): round 4 applies
 * the <= n+1 Givens rotations directly instead of forming the dense (N+1)^2 matrix G and the
 * dense product cat(Q,1)*G' (RbfModel.jl:462).  The CPU baseline timed from this file is
 * therefore FASTER than the reference's own formulation; it is labelled "port".
 *
 * Layout: sites are AoS (site i = sites[i*n .. i*n+n)), matrices column-major, ids 1-based.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_CUBIC 0
#define ORC_INV_MULTIQUADRIC 1
#define ORC_MULTIQUADRIC 2
#define ORC_THIN_PLATE_SPLINE 3
#define ORC_GAUSSIAN 4

typedef struct {
    int32_t kernel;             /* ORC_* */
    int32_t poly_degree;        /* -1, 0, 1 */
    double alpha;               /* shape parameter (resolved; 1 when NaN) */
    double beta;                /* cubic exponent / multiquadric exponent / tps k */
    double theta_enlarge_1, theta_enlarge_2, theta_pivot, theta_pivot_cholesky;
    int32_t max_model_points;   /* <= 0: (n+1)(n+2)/2 */
    int32_t optimized_sampling;
} orc_cfg;

/* rtol of Julia's `isapprox(delta, delta_max)` at RbfModel.jl:588: max(sqrt(eps(T)) over the two argument types) -- sqrt(eps(Float64))
 * when the algorithm config holds Float64 radii, sqrt(eps(Float32)) = 3.4526698e-4 with the default config, whose delta_max is a
 * Float32 literal (AbstractConfigInterface.jl:31).  Process-wide setting of the test infrastructure (set once before a run). */
static double orc_isapprox_rtol = 1.4901161193847656e-08;
void orc_set_isapprox_rtol(double rtol) { orc_isapprox_rtol = rtol; }

/* ---------------------------------------------------------------- radial functions */
static double sgn_pow(int e) { return (e & 1) ? -1.0 : 1.0; }

static double orc_phi(const orc_cfg* c, double rho) {
    switch (c->kernel) {
    case ORC_CUBIC: return sgn_pow((int)ceil(c->beta / 2)) * pow(rho, c->beta);
    case ORC_MULTIQUADRIC: return sgn_pow((int)ceil(c->beta)) * pow(1.0 + (c->alpha * rho) * (c->alpha * rho), c->beta);
    case ORC_INV_MULTIQUADRIC: return pow(1.0 + (c->alpha * rho) * (c->alpha * rho), -c->beta);
    case ORC_GAUSSIAN: return exp(-(c->alpha * rho) * (c->alpha * rho));
    case ORC_THIN_PLATE_SPLINE: {
        int k = (int)c->beta;
        return rho == 0 ? 0.0 : sgn_pow(k + 1) * pow(rho, 2 * k) * log(rho);
    }
    }
    return NAN;
}

/* phi'(rho)/rho, zero contribution at rho == 0 where singular */
static double orc_psi(const orc_cfg* c, double rho) {
    double a = c->alpha, b = c->beta;
    switch (c->kernel) {
    case ORC_CUBIC: return rho == 0 ? 0.0 : sgn_pow((int)ceil(b / 2)) * b * pow(rho, b - 2);
    case ORC_MULTIQUADRIC: return sgn_pow((int)ceil(b)) * 2 * b * a * a * pow(1.0 + (a * rho) * (a * rho), b - 1);
    case ORC_INV_MULTIQUADRIC: return -2 * b * a * a * pow(1.0 + (a * rho) * (a * rho), -b - 1);
    case ORC_GAUSSIAN: return -2 * a * a * exp(-(a * rho) * (a * rho));
    case ORC_THIN_PLATE_SPLINE: {
        int k = (int)b;
        return rho == 0 ? 0.0 : sgn_pow(k + 1) * pow(rho, 2 * k - 2) * (2 * k * log(rho) + 1.0);
    }
    }
    return NAN;
}

static int orc_cpd_order(const orc_cfg* c) {
    switch (c->kernel) {
    case ORC_CUBIC: return (int)ceil(c->beta / 2);
    case ORC_MULTIQUADRIC: return (int)ceil(c->beta);
    case ORC_THIN_PLATE_SPLINE: return (int)c->beta + 1;
    default: return 0;
    }
}

static int poly_dim(int n, int deg) { return deg < 0 ? 0 : (deg == 0 ? 1 : n + 1); }

static double dist2(const double* a, const double* b, int n) {
    double s = 0;
    for (int i = 0; i < n; ++i) { double d = a[i] - b[i]; s += d * d; }
    return s;
}

/* ---------------------------------------------------------------- small dense LA (LAPACK conventions) */
/* dlarfg: x (length m-1) scaled in place, alpha overwritten by beta, returns tau */
static double larfg(int m, double* alpha, double* x) {
    if (m <= 1) return 0.0;
    double xnorm = 0;
    for (int i = 0; i < m - 1; ++i) xnorm = hypot(xnorm, x[i]);
    if (xnorm == 0) return 0.0;
    double beta = -copysign(hypot(*alpha, xnorm), *alpha);
    double tau = (beta - *alpha) / beta;
    double sc = 1.0 / (*alpha - beta);
    for (int i = 0; i < m - 1; ++i) x[i] *= sc;
    *alpha = beta;
    return tau;
}

/* dgeqr2 on A (m x nc, ld lda); tau[min(m,nc)] */
static void geqr2(int m, int nc, double* A, int lda, double* tau) {
    int k = m < nc ? m : nc;
    for (int i = 0; i < k; ++i) {
        tau[i] = larfg(m - i, &A[i + i * lda], &A[(i + 1 < m ? i + 1 : i) + i * lda]);
        if (tau[i] != 0)
            for (int c = i + 1; c < nc; ++c) {
                double w = A[i + c * lda];
                for (int r = i + 1; r < m; ++r) w += A[r + i * lda] * A[r + c * lda];
                w *= tau[i];
                A[i + c * lda] -= w;
                for (int r = i + 1; r < m; ++r) A[r + c * lda] -= w * A[r + i * lda];
            }
    }
}

/* full m x m Q = H_0 ... H_{k-1} from geqr2 output (dorg2r semantics) */
static void orgq_full(int m, int k, const double* A, int lda, const double* tau, double* Q, int ldq) {
    for (int c = 0; c < m; ++c)
        for (int r = 0; r < m; ++r) Q[r + c * ldq] = (r == c) ? 1.0 : 0.0;
    for (int i = k - 1; i >= 0; --i) {
        if (tau[i] == 0) continue;
        for (int c = i; c < m; ++c) {
            double w = Q[i + c * ldq];
            for (int r = i + 1; r < m; ++r) w += A[r + i * lda] * Q[r + c * ldq];
            w *= tau[i];
            Q[i + c * ldq] -= w;
            for (int r = i + 1; r < m; ++r) Q[r + c * ldq] -= w * A[r + i * lda];
        }
    }
}

/* LinearAlgebra.givensAlgorithm convention */
static void givens(double f, double g, double* c, double* s) {
    if (g == 0) { *c = 1; *s = 0; return; }
    if (f == 0) { *c = 0; *s = 1; return; }
    double r = hypot(f, g);
    *c = f / r; *s = g / r;
    if (fabs(f) > fabs(g) && *c < 0) { *c = -*c; *s = -*s; }
}

/* LU with partial pivoting (dgetf2) + solve for nrhs right-hand sides; returns 0 or index of zero pivot+1 */
static int lu_solve(int n, double* A, int lda, double* B, int ldb, int nrhs) {
    int* piv = (int*)malloc(sizeof(int) * (size_t)n);
    int info = 0;
    for (int j = 0; j < n; ++j) {
        int p = j; double mx = fabs(A[j + j * lda]);
        for (int r = j + 1; r < n; ++r) if (fabs(A[r + j * lda]) > mx) { mx = fabs(A[r + j * lda]); p = r; }
        piv[j] = p;
        if (mx == 0) { if (!info) info = j + 1; continue; }
        if (p != j) for (int c = 0; c < n; ++c) { double t = A[j + c * lda]; A[j + c * lda] = A[p + c * lda]; A[p + c * lda] = t; }
        double inv = 1.0 / A[j + j * lda];
        for (int r = j + 1; r < n; ++r) A[r + j * lda] *= inv;
        for (int c = j + 1; c < n; ++c) {
            double a = A[j + c * lda];
            if (a != 0) for (int r = j + 1; r < n; ++r) A[r + c * lda] -= A[r + j * lda] * a;
        }
    }
    if (!info)
        for (int q = 0; q < nrhs; ++q) {
            double* b = B + (size_t)q * ldb;
            for (int j = 0; j < n; ++j) if (piv[j] != j) { double t = b[j]; b[j] = b[piv[j]]; b[piv[j]] = t; }
            for (int j = 0; j < n; ++j) { double a = b[j]; for (int r = j + 1; r < n; ++r) b[r] -= A[r + j * lda] * a; }
            for (int j = n - 1; j >= 0; --j) { b[j] /= A[j + j * lda]; double a = b[j]; for (int r = 0; r < j; ++r) b[r] -= A[r + j * lda] * a; }
        }
    free(piv);
    return info;
}

/* ---------------------------------------------------------------- geometry */
static void local_bounds(int n, const double* x, double delta, const double* glb, const double* gub,
                         double* lb, double* ub) {
    for (int i = 0; i < n; ++i) {
        lb[i] = fmax(glb[i], x[i] - delta);
        ub[i] = fmin(gub[i], x[i] + delta);
    }
}

static int in_box(int n, const double* s, const double* lb, const double* ub) {
    for (int i = 0; i < n; ++i) if (!(lb[i] <= s[i] && s[i] <= ub[i])) return 0;
    return 1;
}

/* intersect_box(...; return_vals = :absmax) */
double orc_intersect_box_absmax(int n, const double* x, const double* d, const double* lb, const double* ub) {
    int any = 0;
    for (int i = 0; i < n; ++i) if (d[i] != 0) any = 1;
    if (!any) return INFINITY;
    double s_pos = INFINITY, s_neg = -INFINITY; int have_pos = 0, have_neg = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (int i = 0; i < n; ++i) {
            if (d[i] == 0) continue;
            double tmp = (pass == 0 ? lb[i] : ub[i]) - x[i];
            double sig;
            if (tmp != 0) sig = tmp / d[i];
            else if (pass == 0) sig = d[i] > 0 ? INFINITY : 0.0;
            else sig = d[i] < 0 ? INFINITY : 0.0;
            if (sig >= 0) { if (!have_pos || sig < s_pos) s_pos = sig; have_pos = 1; }
            else { if (!have_neg || sig > s_neg) s_neg = sig; have_neg = 1; }   /* NaN lands here like !(sig>=0) */
        }
    if (!have_pos) s_pos = 0.0;
    if (!have_neg) s_neg = 0.0;
    return fabs(s_pos) >= fabs(s_neg) ? s_pos : s_neg;
}

/* ---------------------------------------------------------------- affinely independent filter */
/* Z (n x (n-j)) <- trailing columns of the full Q of Y (n x j), columns scaled by inf-norm. */
static void orth_complement(int n, int j, const double* Y, double* Z, double* work /* n*j + n + n*n */) {
    double* A = work; double* tau = A + (size_t)n * j; double* Q = tau + n;
    memcpy(A, Y, sizeof(double) * (size_t)n * j);
    geqr2(n, j, A, n, tau);
    int k = n < j ? n : j;
    orgq_full(n, k, A, n, tau, Q, n);
    for (int c = j; c < n; ++c) {
        double mx = 0;
        for (int r = 0; r < n; ++r) mx = fmax(mx, fabs(Q[r + c * n]));
        for (int r = 0; r < n; ++r) Z[r + (c - j) * n] = Q[r + c * n] / mx;
    }
}

/* One filter run.  cand[0..nc) ascending ids; on return picked[0..*npicked) are ids in selection
 * order.  Y (n x n storage, *jY columns used) and Z (n x n storage) are in/out.  margin: minimum
 * relative gap between the winning score and max(runner-up, pivot) (knife-edge detector). */
static void filter_run(int n, const double* sites, const double* x, const int* cand, int nc, double piv,
                       int n_wanted, double* Y, int* jY, double* Z, int* picked, int* npicked, double* margin) {
    *npicked = 0;
    if (nc == 0) return;
    double* work = (double*)malloc(sizeof(double) * ((size_t)n * n * 2 + 4 * (size_t)n));
    char* used = (char*)calloc((size_t)nc, 1);
    double* s = work + (size_t)n * n * 2 + n;
    double* t = s + n;
    /* first pick: argmax inf-norm of shifted seed, unconditional */
    double best = -INFINITY, second = -INFINITY; int bi = -1;
    for (int c = 0; c < nc; ++c) {
        const double* st = sites + (size_t)(cand[c] - 1) * n;
        double v = 0;
        for (int i = 0; i < n; ++i) v = fmax(v, fabs(st[i] - x[i]));
        if (v > best) { second = best; best = v; bi = c; } else if (v > second) second = v;
    }
    if (bi < 0) bi = 0;                       /* all-NaN corner: findmax returns index 1 */
    (void)second;   /* first pick is exact arithmetic (max |s_i|): ties resolve identically everywhere */
    int found = 0;
    for (;;) {
        const double* st = sites + (size_t)(cand[bi] - 1) * n;
        for (int i = 0; i < n; ++i) Y[i + (size_t)(*jY) * n] = st[i] - x[i];
        (*jY)++;
        orth_complement(n, *jY, Y, Z, work);
        used[bi] = 1; picked[found++] = cand[bi];
        if (found == n_wanted) break;
        int zc = n - *jY;
        best = -INFINITY; second = -INFINITY; bi = -1;
        int remaining = 0;
        for (int c = 0; c < nc; ++c) {
            if (used[c]) continue;
            remaining++;
            const double* sc = sites + (size_t)(cand[c] - 1) * n;
            for (int i = 0; i < n; ++i) s[i] = sc[i] - x[i];
            double v = 0;
            if (zc > 0) {
                for (int q = 0; q < zc; ++q) { double a = 0; for (int i = 0; i < n; ++i) a += Z[i + (size_t)q * n] * s[i]; t[q] = a; }
                for (int i = 0; i < n; ++i) { double a = 0; for (int q = 0; q < zc; ++q) a += Z[i + (size_t)q * n] * t[q]; v = fmax(v, fabs(a)); }
            }
            if (v > best) { second = best; best = v; bi = c; } else if (v > second) second = v;
        }
        if (!remaining) break;
        if (best > 0) *margin = fmin(*margin, (best - fmax(second, piv)) / best >= 0
                                              ? (best - fmax(second, piv)) / best : (piv - best) / piv);
        if (!(best > piv)) break;
    }
    *npicked = found;
    free(used); free(work);
}

/* ---------------------------------------------------------------- round 4 */
/* centers: N0 found-set sites (AoS, N0 x n).  cand: ascending candidate ids (sites from `sites`).
 * Returns accepted ids in r4.  margins[1] <- min |tau2 - thr| / max(|tau2|, thr). */
static int round4_core(const orc_cfg* cfg, int n, const double* sites, const double* centers0, int N0,
                       const int* cand, int nc, int* r4, double* tau_margin) {
    int max_points = cfg->max_model_points <= 0 ? ((n + 1) * (n + 2)) / 2 : cfg->max_model_points;
    int nr4 = 0;
    if (!(N0 < max_points && nc > 0)) return 0;
    int deg = cfg->poly_degree, p = poly_dim(n, deg);
    int NM = max_points > N0 ? max_points : N0; NM += 1;
    int MM = NM;
    double thr = cfg->theta_pivot_cholesky * cfg->theta_pivot_cholesky; thr *= thr;   /* squared twice */
    double* C = (double*)malloc(sizeof(double) * (size_t)NM * n);
    double* Phi = (double*)calloc((size_t)NM * NM, sizeof(double));
    double* Q = (double*)calloc((size_t)NM * NM, sizeof(double));
    double* R = (double*)calloc((size_t)NM * (p > 0 ? p : 1), sizeof(double));
    double* Z = (double*)calloc((size_t)NM * MM, sizeof(double));
    double* Li = (double*)calloc((size_t)MM * MM, sizeof(double));
    double* tau = (double*)calloc((size_t)NM + p + 1, sizeof(double));
    double* vec = (double*)malloc(sizeof(double) * (size_t)(8 * NM + 4 * (p + 1)));
    double *phix = vec, *q = phix + NM, *u = q + NM, *v = u + NM, *t = v + NM, *row = t + NM,
           *cs = row + NM, *sn = cs + (p + 1), *gt = sn + (p + 1), *rl = gt + (p + 1);
    int N = N0, m = 0;
    memcpy(C, centers0, sizeof(double) * (size_t)N0 * n);
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j)
        Phi[i + (size_t)j * NM] = orc_phi(cfg, sqrt(dist2(C + (size_t)i * n, C + (size_t)j * n, n)));
    /* Pi (N x p) -> QR -> full Q (N x N), R (N x p, zero-padded) */
    if (p > 0) {
        double* A = (double*)malloc(sizeof(double) * (size_t)N * p);
        for (int i = 0; i < N; ++i) { A[i] = 1.0; for (int c = 1; c < p; ++c) A[i + (size_t)c * N] = C[(size_t)i * n + c - 1]; }
        geqr2(N, p, A, N, tau);
        int k = N < p ? N : p;
        double* Qf = (double*)malloc(sizeof(double) * (size_t)N * N);
        orgq_full(N, k, A, N, tau, Qf, N);
        for (int c = 0; c < N; ++c) for (int r = 0; r < N; ++r) Q[r + (size_t)c * NM] = Qf[r + (size_t)c * N];
        for (int c = 0; c < p; ++c) for (int r = 0; r < N; ++r) R[r + (size_t)c * NM] = (r <= c) ? A[r + (size_t)c * N] : 0.0;
        free(A); free(Qf);
    } else {
        for (int i = 0; i < N; ++i) Q[i + (size_t)i * NM] = 1.0;
    }
    double phi0 = Phi[0];
    int full_rank_dim = deg < 0 ? 0 : p;     /* binomial(n+deg, n) for deg in {0,1} */
    for (int ci = 0; ci < nc && N < max_points; ++ci) {
        const double* xi = sites + (size_t)(cand[ci] - 1) * n;
        for (int i = 0; i < N; ++i) phix[i] = orc_phi(cfg, sqrt(dist2(xi, C + (size_t)i * n, n)));
        /* Givens on [R; pi_xi'] against pivots j < J */
        int J = N < p ? N : p;
        for (int c = 0; c < p; ++c) rl[c] = (c == 0) ? 1.0 : xi[c - 1];
        for (int j = 0; j < J; ++j) {
            givens(R[j + (size_t)j * NM], rl[j], &cs[j], &sn[j]);
            for (int c = j; c < p; ++c) rl[c] = -sn[j] * R[j + (size_t)c * NM] + cs[j] * rl[c];
        }
        if (N < full_rank_dim) {
            double nr = 0;
            for (int c = 0; c < p; ++c) nr = hypot(nr, rl[c]);
            if (nr <= 2.220446049250313e-16 * 10) continue;
        }
        /* last row of G: g_tilde[j] = -s_j prod_{i>j} c_i ; g_hat = prod c_i */
        double gh = 1.0;
        for (int j = J - 1; j >= 0; --j) { gt[j] = -sn[j] * gh; gh *= cs[j]; }
        for (int i = 0; i < N; ++i) { double a = 0; for (int j = 0; j < J; ++j) a += Q[i + (size_t)j * NM] * gt[j]; q[i] = a; }
        for (int i = 0; i < N; ++i) { double a = 0; for (int j = 0; j < N; ++j) a += Phi[i + (size_t)j * NM] * q[j]; u[i] = a; }
        double sig = 0, pq = 0;
        for (int i = 0; i < N; ++i) { sig += q[i] * u[i]; pq += phix[i] * q[i]; }
        sig += 2 * gh * pq + gh * gh * phi0;
        for (int i = 0; i < N; ++i) u[i] += gh * phix[i];
        for (int c = 0; c < m; ++c) { double a = 0; for (int i = 0; i < N; ++i) a += Z[i + (size_t)c * NM] * u[i]; v[c] = a; }
        double tn = 0;
        for (int r = 0; r < m; ++r) { double a = 0; for (int c = 0; c <= r; ++c) a += Li[r + (size_t)c * MM] * v[c]; t[r] = a; tn += a * a; }
        double nrm = sqrt(tn);
        double tau2 = sig - nrm * nrm;
        double mg = fabs(tau2 - thr) / fmax(fabs(tau2), thr);
        if (mg < *tau_margin) *tau_margin = mg;
        if (tau2 > thr) {
            r4[nr4++] = cand[ci];
            double tv = sqrt(tau2);
            /* Q <- blkdiag(Q,1) * G' : rotate column pairs (j, N) */
            for (int i = 0; i <= N; ++i) Q[i + (size_t)N * NM] = (i == N) ? 1.0 : 0.0;
            for (int j = 0; j < N; ++j) Q[N + (size_t)j * NM] = 0.0;
            for (int j = 0; j < J; ++j)
                for (int i = 0; i <= N; ++i) {
                    double a = Q[i + (size_t)j * NM], b = Q[i + (size_t)N * NM];
                    Q[i + (size_t)j * NM] = cs[j] * a + sn[j] * b;
                    Q[i + (size_t)N * NM] = -sn[j] * a + cs[j] * b;
                }
            /* Z <- [Z q; 0 gh] */
            for (int c = 0; c < m; ++c) Z[N + (size_t)c * NM] = 0.0;
            for (int i = 0; i < N; ++i) Z[i + (size_t)m * NM] = q[i];
            Z[N + (size_t)m * NM] = gh;
            /* Linv <- [Linv 0; -(t' Linv)/tau 1/tau] */
            for (int c = 0; c < m; ++c) { double a = 0; for (int r = c; r < m; ++r) a += t[r] * Li[r + (size_t)c * MM]; row[c] = -a / tv; }
            for (int c = 0; c < m; ++c) { Li[m + (size_t)c * MM] = row[c]; Li[c + (size_t)m * MM] = 0.0; }
            Li[m + (size_t)m * MM] = 1.0 / tv;
            /* R <- rotated [R; pi'] */
            for (int c = 0; c < p; ++c) rl[c] = (c == 0) ? 1.0 : xi[c - 1];
            for (int j = 0; j < J; ++j)
                for (int c = 0; c < p; ++c) {
                    double a = R[j + (size_t)c * NM], b = rl[c];
                    R[j + (size_t)c * NM] = cs[j] * a + sn[j] * b;
                    rl[c] = -sn[j] * a + cs[j] * b;
                }
            for (int c = 0; c < p; ++c) R[N + (size_t)c * NM] = rl[c];
            /* Phi, centres */
            for (int i = 0; i < N; ++i) { Phi[i + (size_t)N * NM] = phix[i]; Phi[N + (size_t)i * NM] = phix[i]; }
            Phi[N + (size_t)N * NM] = phi0;
            memcpy(C + (size_t)N * n, xi, sizeof(double) * (size_t)n);
            N++; m++;
        }
    }
    free(C); free(Phi); free(Q); free(R); free(Z); free(Li); free(tau); free(vec);
    return nr4;
}

/* _rbf_round4 with an explicit found set (ids into sites), as test/rbf_models.jl:74-86 calls it. */
int orc_round4(const orc_cfg* cfg, int n, int n_db, const double* sites, const double* lb2, const double* ub2,
               const int* found, int n_found, int* r4, double* margins) {
    int* cand = (int*)malloc(sizeof(int) * (size_t)(n_db > 0 ? n_db : 1));
    int nc = 0;
    for (int id = 1; id <= n_db; ++id) {
        int ex = 0;
        for (int f = 0; f < n_found; ++f) if (found[f] == id) { ex = 1; break; }
        if (!ex && in_box(n, sites + (size_t)(id - 1) * n, lb2, ub2)) cand[nc++] = id;
    }
    double* C = (double*)malloc(sizeof(double) * (size_t)(n_found > 0 ? n_found : 1) * n);
    for (int f = 0; f < n_found; ++f) memcpy(C + (size_t)f * n, sites + (size_t)(found[f] - 1) * n, sizeof(double) * (size_t)n);
    double tm = INFINITY;
    int r = round4_core(cfg, n, sites, C, n_found, cand, nc, r4, &tm);
    if (margins) margins[1] = tm;
    free(cand); free(C);
    return r;
}

/* prepare_update_model for one instance (no meta_array sharing: that is host logic over ids).
 * Outputs: r1/r2 (<= n ids each), r3_sites (<= n new sites, AoS), r4 (<= max_points ids),
 * dirs (n x n_dirs col-major = improving_directions after round 1 / identity on rebuild),
 * flags_out[0] fully_linear, flags_out[1] rebuilt (round-3 pivot failure -> coordinate rebuild),
 * margins[0] filter decision margin, margins[1] tau^2 margin. */
int orc_select_points(const orc_cfg* cfg, int n, int n_db, const double* sites, int x_index, const double* x,
                      double delta, double delta_max, const double* glb, const double* gub,
                      int ensure_fully_linear, int force_rebuild, int max_new,
                      int* r1, int* n_r1, int* r2, int* n_r2, double* r3_sites, int* n_r3,
                      int* r4, int* n_r4, double* dirs, int* n_dirs, int* flags_out, double* margins) {
    double* lb1 = (double*)malloc(sizeof(double) * 4 * (size_t)n);
    double *ub1 = lb1 + n, *lb2 = ub1 + n, *ub2 = lb2 + n;
    double delta_1 = cfg->theta_enlarge_1 * delta;
    double piv = cfg->theta_pivot * delta_1;
    double delta_2 = cfg->theta_enlarge_2 * delta_max;
    local_bounds(n, x, delta_1, glb, gub, lb1, ub1);
    local_bounds(n, x, delta_2, glb, gub, lb2, ub2);
    double* Y = (double*)calloc((size_t)n * n * 2, sizeof(double));
    double* Z = Y + (size_t)n * n;
    int* cand1 = (int*)malloc(sizeof(int) * (size_t)(n_db > 0 ? 2 * n_db : 2));
    int* cand2 = cand1 + (n_db > 0 ? n_db : 1);
    char* in1 = (char*)calloc((size_t)(n_db > 0 ? n_db : 1), 1);
    double fmargin = INFINITY, tmargin = INFINITY;
    int rebuilt = 0, fully_linear = 0;
    *n_r1 = *n_r2 = *n_r3 = *n_r4 = 0;
restart:
    fully_linear = 0;
    *n_r1 = *n_r2 = *n_r3 = 0;
    int jY = 0, nc1 = 0;
    for (int c = 0; c < n; ++c) for (int r = 0; r < n; ++r) Z[r + (size_t)c * n] = (r == c) ? 1.0 : 0.0;
    int skip_search = force_rebuild || !cfg->optimized_sampling;
    if (skip_search) {
        for (int c = 0; c < n; ++c) for (int r = 0; r < n; ++r) dirs[r + (size_t)c * n] = (r == c) ? 1.0 : 0.0;
        *n_dirs = n;
    } else {
        for (int id = 1; id <= n_db; ++id)
            if (id != x_index && in_box(n, sites + (size_t)(id - 1) * n, lb1, ub1)) { cand1[nc1++] = id; in1[id - 1] = 1; }
        filter_run(n, sites, x, cand1, nc1, piv, n, Y, &jY, Z, r1, n_r1, &fmargin);
        int zc = n - jY;                                   /* improving directions: reverse(eachcol(Z)) */
        for (int c = 0; c < zc; ++c) memcpy(dirs + (size_t)c * n, Z + (size_t)(zc - 1 - c) * n, sizeof(double) * (size_t)n);
        *n_dirs = zc;
    }
    int n_missing = n - *n_r1;
    int approx = fabs(delta - delta_max) <= orc_isapprox_rtol * fmax(fabs(delta), fabs(delta_max));
    if (n_missing == 0 || skip_search || ensure_fully_linear || (approx && cfg->theta_enlarge_1 == cfg->theta_enlarge_2)) {
        fully_linear = 1;
    } else {
        int nc2 = 0;
        for (int id = 1; id <= n_db; ++id)
            if (id != x_index && !in1[id - 1] && in_box(n, sites + (size_t)(id - 1) * n, lb2, ub2)) cand2[nc2++] = id;
        filter_run(n, sites, x, cand2, nc2, piv, n_missing, Y, &jY, Z, r2, n_r2, &fmargin);
    }
    n_missing -= *n_r2;
    if (n_missing > 0) {
        int n_new = n_missing < max_new ? n_missing : max_new;
        if (n_new < 0) n_new = 0;
        int fl = n_new >= n_missing;
        int failed = 0;
        for (int i = 0; i < n_new; ++i) {
            const double* d = dirs + (size_t)i * n;
            double len = orc_intersect_box_absmax(n, x, d, lb1, ub1);
            double on = 0;
            for (int r = 0; r < n; ++r) { double o = len * d[r]; r3_sites[(size_t)i * n + r] = x[r] + o; on = fmax(on, fabs(o)); }
            if (on <= piv) {
                if (ensure_fully_linear && !force_rebuild) { failed = 1; break; }
                fl = 0;
            }
        }
        if (failed) {
            force_rebuild = 1; ensure_fully_linear = 1; rebuilt = 1;
            memset(in1, 0, (size_t)(n_db > 0 ? n_db : 1));
            goto restart;
        }
        *n_r3 = n_new;
        fully_linear = fl && (*n_r2 == 0);
    }
    if (cfg->optimized_sampling) {
        /* found set: centre, r1, r2, r3 (new sites); candidates: box 2 minus found ids */
        int nf = 1 + *n_r1 + *n_r2 + *n_r3;
        double* C = (double*)malloc(sizeof(double) * (size_t)nf * n);
        int f = 0;
        memcpy(C, sites + (size_t)(x_index - 1) * n, sizeof(double) * (size_t)n); f++;
        for (int i = 0; i < *n_r1; ++i) memcpy(C + (size_t)(f++) * n, sites + (size_t)(r1[i] - 1) * n, sizeof(double) * (size_t)n);
        for (int i = 0; i < *n_r2; ++i) memcpy(C + (size_t)(f++) * n, sites + (size_t)(r2[i] - 1) * n, sizeof(double) * (size_t)n);
        for (int i = 0; i < *n_r3; ++i) memcpy(C + (size_t)(f++) * n, r3_sites + (size_t)i * n, sizeof(double) * (size_t)n);
        int nc4 = 0;
        int* cand4 = cand1;
        for (int id = 1; id <= n_db; ++id) {
            if (id == x_index) continue;
            int ex = 0;
            for (int i = 0; i < *n_r1 && !ex; ++i) ex = (r1[i] == id);
            for (int i = 0; i < *n_r2 && !ex; ++i) ex = (r2[i] == id);
            if (!ex && in_box(n, sites + (size_t)(id - 1) * n, lb2, ub2)) cand4[nc4++] = id;
        }
        *n_r4 = round4_core(cfg, n, sites, C, nf, cand4, nc4, r4, &tmargin);
        free(C);
    }
    flags_out[0] = fully_linear; flags_out[1] = rebuilt;
    if (margins) { margins[0] = fmargin; margins[1] = tmargin; }
    free(lb1); free(Y); free(cand1); free(in1);
    return 0;
}

/* ---------------------------------------------------------------- model build (saddle system, LU) */
/* sites N x n (AoS), values N x k (row i = values of site i).  Outputs w (N x k, row-major like
 * values) and lam (p x k row-major), p = poly_dim(n, max(deg, cpd-1)).  Returns 0, or >0 singular.
 * N < p: minimum-norm lam with w = 0 (assumption U9). */
int orc_build(const orc_cfg* cfg, int n, int k, int N, const double* sites, const double* values,
              double* w, double* lam) {
    int deg = cfg->poly_degree; int cpd = orc_cpd_order(cfg);
    if (deg < cpd - 1) deg = cpd - 1;
    if (deg > 1) return -1;
    int p = poly_dim(n, deg);
    if (N < p) {
        /* w = 0;  Pi lam = Y minimum norm:  lam = Pi' (Pi Pi')^{-1} Y */
        double* G = (double*)malloc(sizeof(double) * (size_t)N * N);
        double* B = (double*)malloc(sizeof(double) * (size_t)N * k);
        for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j) {
            double a = 1.0;
            if (p > 1) for (int c = 0; c < n; ++c) a += sites[(size_t)i * n + c] * sites[(size_t)j * n + c];
            G[i + (size_t)j * N] = a;
        }
        for (int q = 0; q < k; ++q) for (int i = 0; i < N; ++i) B[i + (size_t)q * N] = values[(size_t)i * k + q];
        int info = lu_solve(N, G, N, B, N, k);
        for (int i = 0; i < N * k; ++i) w[i] = 0.0;
        for (int c = 0; c < p; ++c) for (int q = 0; q < k; ++q) {
            double a = 0;
            for (int i = 0; i < N; ++i) a += ((c == 0) ? 1.0 : sites[(size_t)i * n + c - 1]) * B[i + (size_t)q * N];
            lam[(size_t)c * k + q] = a;
        }
        free(G); free(B);
        return info;
    }
    int S = N + p;
    double* A = (double*)calloc((size_t)S * S, sizeof(double));
    double* B = (double*)calloc((size_t)S * k, sizeof(double));
    for (int i = 0; i < N; ++i) for (int j = 0; j < N; ++j)
        A[i + (size_t)j * S] = orc_phi(cfg, sqrt(dist2(sites + (size_t)i * n, sites + (size_t)j * n, n)));
    for (int i = 0; i < N; ++i) for (int c = 0; c < p; ++c) {
        double v = (c == 0) ? 1.0 : sites[(size_t)i * n + c - 1];
        A[i + (size_t)(N + c) * S] = v; A[(N + c) + (size_t)i * S] = v;
    }
    for (int q = 0; q < k; ++q) for (int i = 0; i < N; ++i) B[i + (size_t)q * S] = values[(size_t)i * k + q];
    int info = lu_solve(S, A, S, B, S, k);
    for (int i = 0; i < N; ++i) for (int q = 0; q < k; ++q) w[(size_t)i * k + q] = B[i + (size_t)q * S];
    for (int c = 0; c < p; ++c) for (int q = 0; q < k; ++q) lam[(size_t)c * k + q] = B[N + c + (size_t)q * S];
    free(A); free(B);
    return info;
}

/* ---------------------------------------------------------------- evaluation, one point at a time */
/* centers N x n AoS; w N x k; lam p x k; X M x n AoS; Y M x k; Jout M x k x n (row-major) */
static int eff_degree(const orc_cfg* cfg) { int d = cfg->poly_degree, c = orc_cpd_order(cfg); return d < c - 1 ? c - 1 : d; }

void orc_eval(const orc_cfg* cfg, int n, int k, int N, const double* centers, const double* w, const double* lam,
              long M, const double* X, double* Y, int nthreads) {
    int deg = eff_degree(cfg);
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (long mi = 0; mi < M; ++mi) {
        const double* x = X + (size_t)mi * n;
        double* y = Y + (size_t)mi * k;
        for (int q = 0; q < k; ++q) y[q] = 0;
        for (int i = 0; i < N; ++i) {
            double ph = orc_phi(cfg, sqrt(dist2(x, centers + (size_t)i * n, n)));
            for (int q = 0; q < k; ++q) y[q] += w[(size_t)i * k + q] * ph;
        }
        if (deg >= 0) for (int q = 0; q < k; ++q) y[q] += lam[q];
        if (deg >= 1) for (int c = 0; c < n; ++c) for (int q = 0; q < k; ++q) y[q] += lam[(size_t)(c + 1) * k + q] * x[c];
    }
}

void orc_jac(const orc_cfg* cfg, int n, int k, int N, const double* centers, const double* w, const double* lam,
             long M, const double* X, double* Jout, int nthreads) {
    int deg = eff_degree(cfg);
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (long mi = 0; mi < M; ++mi) {
        const double* x = X + (size_t)mi * n;
        double* J = Jout + (size_t)mi * k * n;
        for (int i = 0; i < k * n; ++i) J[i] = 0;
        for (int i = 0; i < N; ++i) {
            const double* c = centers + (size_t)i * n;
            double ps = orc_psi(cfg, sqrt(dist2(x, c, n)));
            for (int q = 0; q < k; ++q) {
                double g = w[(size_t)i * k + q] * ps;
                for (int r = 0; r < n; ++r) J[(size_t)q * n + r] += g * (x[r] - c[r]);
            }
        }
        if (deg >= 1) for (int q = 0; q < k; ++q) for (int r = 0; r < n; ++r) J[(size_t)q * n + r] += lam[(size_t)(r + 1) * k + q];
    }
}

/* ---------------------------------------------------------------- batched drivers (one instance per thread,
 * the stand-in for Threads.@threads over optimize() runs, examples/large_scale_benchmarks.jl:253) */
int orc_select_points_batched(const orc_cfg* cfg, int B, int n, int n_db, const double* sites /* B x n_db x n */,
                              const int* x_index, const double* x /* B x n */, const double* delta, double delta_max,
                              const double* glb, const double* gub, const int* flags_in /* B x 2 */, const int* max_new,
                              int* r1, int* n_r1, int* r2, int* n_r2, double* r3_sites, int* n_r3,
                              int r4_stride, int* r4, int* n_r4, double* dirs, int* n_dirs, int* flags_out, double* margins,
                              int nthreads) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1)
    for (int b = 0; b < B; ++b)
        orc_select_points(cfg, n, n_db, sites + (size_t)b * n_db * n, x_index[b], x + (size_t)b * n, delta[b], delta_max,
                          glb, gub, flags_in[2 * b], flags_in[2 * b + 1], max_new[b],
                          r1 + (size_t)b * n, n_r1 + b, r2 + (size_t)b * n, n_r2 + b, r3_sites + (size_t)b * n * n, n_r3 + b,
                          r4 + (size_t)b * r4_stride, n_r4 + b, dirs + (size_t)b * n * n, n_dirs + b, flags_out + 2 * b,
                          margins ? margins + 2 * b : 0);
    return 0;
}

int orc_build_batched(const orc_cfg* cfg, int B, int n, int k, int N_stride, const int* N, const double* sites,
                      const double* values, double* w, double* lam, int* status, int nthreads) {
    int p = poly_dim(n, eff_degree(cfg));
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : 1)
    for (int b = 0; b < B; ++b)
        status[b] = orc_build(cfg, n, k, N[b], sites + (size_t)b * N_stride * n, values + (size_t)b * N_stride * k,
                              w + (size_t)b * N_stride * k, lam + (size_t)b * p * k);
    return 0;
}

/* Built-in objectives of the BASELINE configs (the reference evaluates user functions in the thread that runs optimize()).
 * func_id 1: ZDT3, 2: ZDT1 (both k = 2, inputs clipped to [0,1] like synthetic.py), 3: two parabolas (k = 2). */
static void orc_objective(int func_id, int n, const double* x, double* y) {
    if (func_id == 3) {
        double a = 0, b = 0;
        for (int i = 0; i < n; ++i) { a += (x[i] - 1.0) * (x[i] - 1.0); b += (x[i] + 1.0) * (x[i] + 1.0); }
        y[0] = a; y[1] = b;
        return;
    }
    double f1 = fmin(fmax(x[0], 0.0), 1.0), s = 0;
    for (int i = 1; i < n; ++i) s += fmin(fmax(x[i], 0.0), 1.0);
    double g = 1.0 + 9.0 * (s / (double)(n - 1));
    y[0] = f1;
    if (func_id == 2) y[1] = g * (1.0 - sqrt(f1 / g));
    else y[1] = g * (1.0 - sqrt(f1 / g) - (f1 / g) * sin(10.0 * 3.14159265358979323846 * f1));
}

/* One whole model build per instance and thread, nothing between the two halves but the training-set gather
 * (prepare_update_model -> eval_missing! of the new round-3 sites -> update_model; RbfModel.jl:518-655, 743-767,
 * Databases.jl:258-277): rounds 1-4, values of [centre; r1; r2; r3; r4] read from the database arrays (round-3 sites are
 * evaluated with the built-in objective `func_id`, or left 0 when func_id == 0), saddle-system solve.
 * sites B x n_db x n, values B x n_db x k.  Outputs: N (B), ids (B x ids_stride: training ids, 0 for round-3 sites),
 * w (B x w_stride x k), lam (B x p x k), status (B).  Returns 0. */
int orc_select_and_build_batched(const orc_cfg* cfg, int B, int n, int k, int n_db, const double* sites, const double* values,
                                 const int* x_index, const double* x, const double* delta, double delta_max,
                                 const double* glb, const double* gub, const int* flags_in, const int* max_new, int func_id,
                                 int ids_stride, int* N, int* ids, int w_stride, double* w, double* lam, int* status, int nthreads) {
    const int p = poly_dim(n, eff_degree(cfg));
    const int mp = cfg->max_model_points <= 0 ? ((n + 1) * (n + 2)) / 2 : cfg->max_model_points;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        int* r1 = (int*)malloc(sizeof(int) * (size_t)(2 * n + mp + 8));
        int *r2 = r1 + n, *r4 = r2 + n;
        double* r3 = (double*)malloc(sizeof(double) * ((size_t)n * n * 2 + (size_t)(mp + n + 1) * (n + k)));
        double* dirs = r3 + (size_t)n * n;
        double* S = dirs + (size_t)n * n;
        double* V = S + (size_t)(mp + n + 1) * n;
#pragma omp for schedule(dynamic, 1)
        for (int b = 0; b < B; ++b) {
            const double* sb = sites + (size_t)b * n_db * n;
            const double* vb = values + (size_t)b * n_db * k;
            int n1, n2, n3, n4, nd, fl[2];
            orc_select_points(cfg, n, n_db, sb, x_index[b], x + (size_t)b * n, delta[b], delta_max, glb, gub,
                              flags_in[2 * b], flags_in[2 * b + 1], max_new[b], r1, &n1, r2, &n2, r3, &n3, r4, &n4, dirs, &nd, fl, 0);
            int f = 0;
            int* idb = ids + (size_t)b * ids_stride;
#define ORC_TAKE(id) do { memcpy(S + (size_t)f * n, sb + (size_t)((id) - 1) * n, sizeof(double) * (size_t)n); \
                          memcpy(V + (size_t)f * k, vb + (size_t)((id) - 1) * k, sizeof(double) * (size_t)k); \
                          if (f < ids_stride) idb[f] = (id); ++f; } while (0)
            ORC_TAKE(x_index[b]);
            for (int i = 0; i < n1; ++i) ORC_TAKE(r1[i]);
            for (int i = 0; i < n2; ++i) ORC_TAKE(r2[i]);
            for (int i = 0; i < n3; ++i) {
                memcpy(S + (size_t)f * n, r3 + (size_t)i * n, sizeof(double) * (size_t)n);
                if (func_id > 0 && k == 2) orc_objective(func_id, n, r3 + (size_t)i * n, V + (size_t)f * k);
                else memset(V + (size_t)f * k, 0, sizeof(double) * (size_t)k);
                if (f < ids_stride) idb[f] = 0;
                ++f;
            }
            for (int i = 0; i < n4; ++i) ORC_TAKE(r4[i]);
#undef ORC_TAKE
            N[b] = f;
            if (f <= w_stride)
                status[b] = orc_build(cfg, n, k, f, S, V, w + (size_t)b * w_stride * k, lam + (size_t)b * p * k);
            else status[b] = -1;
        }
        free(r1); free(r3);
    }
    return 0;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
