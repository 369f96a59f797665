// Batched RBF model build, one CTA per (instance, group) system.
//
// Replaces update_model -> RBF.RBFInterpolationModel (src/models/RbfModel.jl:743-767): solve
//     [Phi Pi; Pi' 0] [w; lambda] = [Y; 0]        for k right-hand sides,
// Phi_ij = phi(||x_i - x_j||), Pi = [1 x] (polynomial tail of degree <= 1).
//
// Route (north star): null-space method instead of the reference's dense LU of the (N+p)^2 saddle matrix.
//   1. Householder QR of Pi (N x p), LAPACK geqr2 conventions;
//   2. the same reflectors applied two-sidedly to Phi:  Phi~ = Q' Phi Q  (symmetric rank-2 updates,
//      ~8 N^2 p flop, instead of forming Z explicitly and two N^2 m GEMMs);
//   3. Cholesky of the trailing m x m block  Z' Phi Z  (m = N - p; positive definite for the
//      conditionally positive definite kernels with the signs of mrbf_common.cuh);
//   4. triangular solves for u, back-substitution with R for lambda, w = Q [0; u].
// Mathematically the unique solution of the saddle system, so values/Jacobians agree with the reference to
// O(cond * eps); see DESIGN.md for the measured deltas.
//
// The whole system lives in shared memory when (N^2 + N (p + k)) doubles fit, else in an L2-resident
// global workspace; the code is identical, only the base pointer differs.
#include "mrbf_common.cuh"
#include <cstdlib>
#include "mrbf_kernels.h"

namespace mrbf {

// Right-looking Cholesky (lower) of the m x m block of A starting at (off, off); returns 0, or the failing column + 1 (the same
// value in every thread).  ONE barrier per column: every thread reads the pivot d^2 itself (no thread-0 section), the trailing update
// uses the unscaled column (a_il -= u_i u_l / d^2), and the scaling of column c (u / d, diagonal <- d) is deferred into the phase
// of column c + 1, where nobody reads that column any more.
__device__ int chol_lower(double* A, int ld, int off, int m) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    double prev_d2 = 1.0;
    for (int c = 0; c < m; ++c) {
        const int pc = off + c;
        const double d2 = A[pc + (size_t)pc * ld];           // final: the barrier below closed the update of column c - 1
        if (!(d2 > 0.0)) return c + 1;                        // uniform: every thread read the same value
        const double rd2 = 1.0 / d2;
        const int rem = m - c - 1;
        if (c > 0) {                                          // deferred scaling of column c - 1
            const double dp = sqrt(prev_d2), rdp = 1.0 / dp;
            double* pcol = A + (size_t)(pc - 1) * ld + pc;
            for (int i = tid; i <= rem; i += nt) pcol[i] *= rdp;
            if (tid == 0) A[pc - 1 + (size_t)(pc - 1) * ld] = dp;
        }
        {   // trailing update, lower triangle only: a warp per column, lanes down the rows (coalesced, no index divisions)
            const double* colp = A + (size_t)pc * ld + pc + 1;
            for (int l = warp; l < rem; l += nwarps) {
                const double al = colp[l] * rd2;
                double* dst = A + (size_t)(pc + 1 + l) * ld + pc + 1;
                int i = l + lane;
                for (; i + 96 < rem; i += 128) {             // four independent load pairs in flight per lane
                    const double c0 = colp[i], c1 = colp[i + 32], c2 = colp[i + 64], c3 = colp[i + 96];
                    const double d0 = dst[i], d1 = dst[i + 32], d2_ = dst[i + 64], d3 = dst[i + 96];
                    dst[i] = fma(-c0, al, d0); dst[i + 32] = fma(-c1, al, d1); dst[i + 64] = fma(-c2, al, d2_); dst[i + 96] = fma(-c3, al, d3);
                }
                for (; i < rem; i += 32) dst[i] = fma(-colp[i], al, dst[i]);
            }
        }
        prev_d2 = d2;
        __syncthreads();
    }
    if (m > 0 && tid == 0) { const int pl_ = off + m - 1; A[pl_ + (size_t)pl_ * ld] = sqrt(prev_d2); }
    __syncthreads();
    return 0;
}

// Solve L L' U = Y in place for the k columns of Yv (rows off..off+m), L from chol_lower.  A warp per right-hand side, no block
// barriers inside: forward substitution as column axpys (the column of L is contiguous), backward substitution as column dots.
__device__ void chol_solve(const double* A, int ld, int off, int m, double* Yv, int ldy, int k) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    for (int q = warp; q < k; q += nwarps) {
        double* y = Yv + (size_t)q * ldy;
        for (int c = 0; c < m; ++c) {                 // forward
            const int pc = off + c;
            const double* col = A + (size_t)pc * ld;
            const double yc = y[pc] / col[pc];
            __syncwarp();
            if (lane == 0) y[pc] = yc;
            for (int i = pc + 1 + lane; i < off + m; i += 32) y[i] = fma(-col[i], yc, y[i]);
            __syncwarp();
        }
        for (int c = m - 1; c >= 0; --c) {            // backward with L'
            const int pc = off + c;
            const double* col = A + (size_t)pc * ld;
            double a = 0.0;
            for (int i = pc + 1 + lane; i < off + m; i += 32) a = fma(col[i], y[i], a);
            a = warp_sum(a);
            __syncwarp();
            if (lane == 0) y[pc] = (y[pc] - a) / col[pc];
            __syncwarp();
        }
    }
    __syncthreads();
}


// The same solve with the whole CTA per step and block barriers: better when A lives in the global workspace (every step is an
// L2 round trip, which 32 warps overlap and one warp does not).
__device__ void chol_solve_block(const double* A, int ld, int off, int m, double* Yv, int ldy, int k) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int c = 0; c < m; ++c) {                 // forward
        const int pc = off + c;
        if (tid < k) Yv[pc + (size_t)tid * ldy] /= A[pc + (size_t)pc * ld];
        __syncthreads();
        const int rem = m - c - 1;
        for (int e = tid; e < rem * k; e += nt) {
            const int i = pc + 1 + e % rem, q = e / rem;
            Yv[i + (size_t)q * ldy] = fma(-A[i + (size_t)pc * ld], Yv[pc + (size_t)q * ldy], Yv[i + (size_t)q * ldy]);
        }
        __syncthreads();
    }
    for (int c = m - 1; c >= 0; --c) {            // backward with L'
        const int pc = off + c;
        if (tid < k) Yv[pc + (size_t)tid * ldy] /= A[pc + (size_t)pc * ld];
        __syncthreads();
        for (int e = tid; e < c * k; e += nt) {
            const int l = off + e % c, q = e / c;
            Yv[l + (size_t)q * ldy] = fma(-A[pc + (size_t)l * ld], Yv[pc + (size_t)q * ldy], Yv[l + (size_t)q * ldy]);
        }
        __syncthreads();
    }
}

// NT = 256 for batches (one CTA per system, many systems per SM-wave); NT = 1024 when only a few large systems are built and
// the L2 latency of the global workspace, not the number of CTAs, limits the time.
template <int NT>
__global__ void __launch_bounds__(NT) build_kernel(BuildParams P) {
    extern __shared__ double smem[];
    const int b = blockIdx.x, n = P.n, k = P.k, p = P.p;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int pl = p > 0 ? p : 1;
    const int N = P.N[b];
    // per-instance leading dimension: the system lives in shared memory whenever THIS instance fits
    const int ld = (N > 0 && N <= P.train_stride) ? (N | 1) : 1;
    double* v = smem;                // P.ld
    double* pv = v + P.ld;           // P.ld
    double* tauv = pv + P.ld;        // pl
    double* red = tauv + pl;         // 80
    double* mats = red + 80;
    const bool in_smem = P.ws_in_smem || ((size_t)ld * ld + (size_t)ld * (pl + k) <= (size_t)P.smem_ws_doubles);
    double* ws = in_smem ? mats : P.ws + (size_t)b * P.ws_stride;
    double* A = ws;                              // ld x ld
    double* Pm = A + (size_t)ld * ld;            // ld x pl
    double* Yv = Pm + (size_t)ld * pl;           // ld x k
    double* w_out = P.w + (size_t)b * P.train_stride * k;
    double* lam_out = P.lam + (size_t)b * pl * k;
    if (P.skip && P.skip[b]) return;                         // already built from the kept round-4 factorisation
    if (N <= 0 || N > P.train_stride) { if (tid == 0) P.status[b] = -1; return; }
    const double* sites = P.sites + (size_t)b * P.train_stride * n;
    if (P.centers_out) {
        double* co = P.centers_out + (size_t)b * P.train_stride * n;
        for (int e = tid; e < N * n; e += nt) co[e] = sites[e];
        if (tid == 0) P.N_out[b] = N;
    }
    const double* values = P.values + (size_t)b * P.train_stride * k;
    RadFn rf; rf.kernel = P.kernel; rf.ibeta = P.ibeta; rf.sgn = P.sgn;
    {
        double a = P.alpha_default;
        if (P.shape) { double sp = P.shape[b]; if (sp == sp) a = sp; }      // NaN => default
        rf.alpha2 = a * a;
        if (tid == 0) P.alpha2_out[b] = rf.alpha2;
    }
    if (tid == 0) red[74] = 0.0;
    // ---- assembly (Gram matrix, polynomial block, right-hand sides)
    {
        // when the system itself lives in the global workspace, the (otherwise idle) shared memory stages the sites: the
        // row-per-thread reads of the distance loop are uncoalesced in global memory
        const double* sp = sites;
        if (!in_smem && (size_t)N * n <= (size_t)P.smem_ws_doubles) {
            for (int e = tid; e < N * n; e += nt) mats[e] = sites[e];
            __syncthreads();
            sp = mats;
        }
        if (in_smem && P.stage_off > 0) {
            // sites staged coordinate-major (St[c * Np + i]): lanes down the rows read consecutive doubles, the column's site is a broadcast
            double* St = smem + P.stage_off;
            const int Np = N | 1;
            for (int e = tid; e < N * n; e += nt) { const int i = e / n, c = e % n; St[c * Np + i] = sites[e]; }
            __syncthreads();
            for (int j = warp; j < N; j += nwarps)
                for (int i = j + lane; i < N; i += 32) {
                    double r2 = 0.0;
#pragma unroll 6
                    for (int c = 0; c < n; ++c) { const double d = St[c * Np + i] - St[c * Np + j]; r2 = fma(d, d, r2); }
                    const double ph = rad_phi(rf, r2);
                    A[i + (size_t)j * ld] = ph;
                    A[j + (size_t)i * ld] = ph;
                }
        } else
        for (int j = warp; j < N; j += nwarps) {             // a warp per column, lanes down the rows: lower triangle, then mirror
            const double* sj = sp + (size_t)j * n;
            for (int i = j + lane; i < N; i += 32) {
                const double* si = sp + (size_t)i * n;
                double r2 = 0.0;
                for (int c = 0; c < n; ++c) { const double d = si[c] - sj[c]; r2 = fma(d, d, r2); }
                const double ph = rad_phi(rf, r2);
                A[i + (size_t)j * ld] = ph;
                A[j + (size_t)i * ld] = ph;
            }
        }
        __syncthreads();                                      // the staging area may be reused below
    }
    for (int e = tid; e < N * p; e += nt) { const int i = e % N, c = e / N; Pm[i + (size_t)c * ld] = poly_basis_at(sites + (size_t)i * n, n, c); }
    for (int e = tid; e < N * k; e += nt) { const int i = e % N, q = e / N; Yv[i + (size_t)q * ld] = values[(size_t)i * k + q]; }
    __syncthreads();

    if (N < p) {
        // fewer sites than polynomial basis functions: w = 0, minimum-norm lambda = Pi' (Pi Pi')^{-1} Y  (U9)
        for (int e = tid; e < N * N; e += nt) {
            const int i = e % N, j = e / N;
            double a = 0.0;
            for (int c = 0; c < p; ++c) a = fma(Pm[i + (size_t)c * ld], Pm[j + (size_t)c * ld], a);
            A[i + (size_t)j * ld] = a;
        }
        __syncthreads();
        const int bad0 = chol_lower(A, ld, 0, N);
        if (bad0) { if (tid == 0) P.status[b] = bad0; return; }
        chol_solve(A, ld, 0, N, Yv, ld, k);
        for (int e = tid; e < p * k; e += nt) {
            const int c = e / k, q = e % k;
            double a = 0.0;
            for (int i = 0; i < N; ++i) a = fma(Pm[i + (size_t)c * ld], Yv[i + (size_t)q * ld], a);
            lam_out[(size_t)c * k + q] = a;
        }
        for (int e = tid; e < N * k; e += nt) w_out[e] = 0.0;
        if (tid == 0) P.status[b] = 0;
        return;
    }

    // ---- Householder QR of Pi, reflectors applied to Pi, Y and two-sidedly to Phi.  Three barriers per reflector: the column norm
    // and v' pv are computed by every warp for itself (identical arithmetic, no block reduction), pv - K v is folded into the
    // rank-2 update, and the rows of pv = tau Phi v are split over G threads each.
    const int G = (nt >= 1024) ? 8 : ((nt >= 512) ? 4 : 2);
    for (int j = 0; j < p; ++j) {
        double part = 0.0;
        for (int i = j + 1 + lane; i < N; i += 32) { const double a = Pm[i + (size_t)j * ld]; part = fma(a, a, part); }
        const double xn2 = warp_sum(part);
        const double alpha = Pm[j + (size_t)j * ld];
        double tau = 0.0, sc = 0.0, beta = alpha;
        {
            const double xnorm = sqrt(xn2);
            if (xnorm != 0.0) {
                beta = -copysign(hypot(alpha, xnorm), alpha);
                tau = (beta - alpha) / beta; sc = 1.0 / (alpha - beta);
            }
        }
        for (int i = tid; i < N; i += nt) v[i] = (i < j) ? 0.0 : ((i == j) ? 1.0 : Pm[i + (size_t)j * ld] * sc);
        __syncthreads();                                     // v complete; nobody reads column j of Pi any more in this step
        for (int i = j + 1 + tid; i < N; i += nt) Pm[i + (size_t)j * ld] = v[i];
        if (tid == 0) { Pm[j + (size_t)j * ld] = beta; tauv[j] = tau; }
        if (tau == 0.0) { __syncthreads(); continue; }
        // trailing columns of Pi and all columns of Y: warp per column
        const int ncols = (p - j - 1) + k;
        for (int cc = warp; cc < ncols; cc += nwarps) {
            double* col = (cc < p - j - 1) ? Pm + (size_t)(j + 1 + cc) * ld : Yv + (size_t)(cc - (p - j - 1)) * ld;
            double a = 0.0;
            for (int i = j + lane; i < N; i += 32) a = fma(v[i], col[i], a);
            a = warp_sum(a) * tau;
            for (int i = j + lane; i < N; i += 32) col[i] = fma(-a, v[i], col[i]);
        }
        // pv = tau Phi v (v is zero above row j)
        for (int r0 = 0; r0 < N; r0 += nt / G) {
            const int i = r0 + tid / G, g = tid % G;
            double a0 = 0.0, a1 = 0.0;
            if (i < N) {
                int l = j + g;
                for (; l + G < N; l += 2 * G) { a0 = fma(A[i + (size_t)l * ld], v[l], a0); a1 = fma(A[i + (size_t)(l + G) * ld], v[l + G], a1); }
                if (l < N) a0 = fma(A[i + (size_t)l * ld], v[l], a0);
            }
            a0 += a1;
            for (int o = G >> 1; o > 0; o >>= 1) a0 += __shfl_xor_sync(0xffffffffu, a0, o);
            if (i < N && g == 0) pv[i] = tau * a0;
        }
        __syncthreads();                                     // pv complete
        double part2 = 0.0;
        for (int i = j + lane; i < N; i += 32) part2 = fma(v[i], pv[i], part2);
        const double K = 0.5 * tau * warp_sum(part2);
        // Phi <- H Phi H :  wv = pv - K v ;  Phi -= v wv' + wv v'   (a warp per column; v is zero above row j)
        for (int l = warp; l < N; l += nwarps) {
            const double vl = v[l], wl = fma(-K, vl, pv[l]);
            double* dst = A + (size_t)l * ld;
            int i = ((l >= j) ? 0 : j) + lane;
            for (; i + 96 < N; i += 128) {
                const double d0 = dst[i], d1 = dst[i + 32], d2 = dst[i + 64], d3 = dst[i + 96];
                const double v0 = v[i], v1 = v[i + 32], v2 = v[i + 64], v3 = v[i + 96];
                dst[i] = d0 - (v0 * wl + fma(-K, v0, pv[i]) * vl); dst[i + 32] = d1 - (v1 * wl + fma(-K, v1, pv[i + 32]) * vl);
                dst[i + 64] = d2 - (v2 * wl + fma(-K, v2, pv[i + 64]) * vl); dst[i + 96] = d3 - (v3 * wl + fma(-K, v3, pv[i + 96]) * vl);
            }
            for (; i < N; i += 32) { const double vi = v[i]; dst[i] -= vi * wl + fma(-K, vi, pv[i]) * vl; }
        }
        __syncthreads();
    }
    // ---- Cholesky of Z' Phi Z and the solves
    const int m = N - p;
    const int bad = chol_lower(A, ld, p, m);
    if (bad) { if (tid == 0) P.status[b] = bad; return; }
    if (in_smem) chol_solve(A, ld, p, m, Yv, ld, k); else chol_solve_block(A, ld, p, m, Yv, ld, k);    // rows p.. of Yv now hold u
    for (int o = warp; o < p * k; o += nwarps) {             // top block: (Q'y)_top - Phi~[top, bottom] u, a warp per entry
        const int r = o % p, q = o / p;
        double a = 0.0;
        for (int c = lane; c < m; c += 32) a = fma(A[r + (size_t)(p + c) * ld], Yv[p + c + (size_t)q * ld], a);
        a = warp_sum(a);
        if (lane == 0) Yv[r + (size_t)q * ld] -= a;
    }
    __syncthreads();
    for (int q = warp; q < k; q += nwarps) {                 // R lambda = top: a warp per right-hand side, column axpys (no block barriers)
        double* y = Yv + (size_t)q * ld;
        for (int c = p - 1; c >= 0; --c) {
            const double* col = Pm + (size_t)c * ld;
            const double yc = y[c] / col[c];
            __syncwarp();
            if (lane == 0) y[c] = yc;
            for (int r = lane; r < c; r += 32) y[r] = fma(-col[r], yc, y[r]);
            __syncwarp();
        }
    }
    __syncthreads();
    for (int e = tid; e < p * k; e += nt) { const int c = e / k, q = e % k; lam_out[(size_t)c * k + q] = Yv[c + (size_t)q * ld]; }
    __syncthreads();
    for (int e = tid; e < p * k; e += nt) { const int r = e % p, q = e / p; Yv[r + (size_t)q * ld] = 0.0; }
    __syncthreads();
    for (int q = warp; q < k; q += nwarps) {                 // w = H_0 ... H_{p-1} [0; u], warp per right-hand side
        double* col = Yv + (size_t)q * ld;
        for (int j = p - 1; j >= 0; --j) {
            const double tau = tauv[j];
            if (tau == 0.0) continue;
            const double* vj = Pm + (size_t)j * ld;
            double a = 0.0;
            for (int i = j + 1 + lane; i < N; i += 32) a = fma(vj[i], col[i], a);
            a = (warp_sum(a) + col[j]) * tau;
            for (int i = j + 1 + lane; i < N; i += 32) col[i] = fma(-a, vj[i], col[i]);
            __syncwarp();
            if (lane == 0) col[j] -= a;
            __syncwarp();
        }
    }
    __syncthreads();
    for (int e = tid; e < N * k; e += nt) { const int i = e / k, q = e % k; w_out[e] = Yv[i + (size_t)q * ld]; }
    if (tid == 0) P.status[b] = 0;
}


// ------------------------------------------------------------------------------------------------
// Build from the factorisation kept by round 4 (mrbf_build_prepared).
//
// Round 4 already holds, for the training set [S0; accepted points] in _collect_indices order, the inverse
// Cholesky factor L^{-1} of the kernel matrix reduced to the null space of Pi' (basis n_eta = e_eta - sum_s c_eta[s] e_s).
// The reference notes that reusing this would save work (RbfModel.jl:657-660, "we do not store the matrices calculated in
// round 4"); here it turns the O(N^3) build into two triangular mat-vecs:
//     r = Y_acc - C' Y_0,   u = L^{-T} L^{-1} r,   w = [-C u; u],   lambda~ = Pi_0^{-1} (Y_0 - G u)
// (G = Phi(S0, acc) - Phi00 C).  lambda~ lives in the centred/scaled basis of round 4 and is mapped back to (1, x).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) build_prepared_kernel(PreparedBuildParams P) {
    extern __shared__ double smem[];
    const int b = blockIdx.x, n = P.n, k = P.k, p = P.p, NM = P.NM, tid = threadIdx.x, nt = blockDim.x;
    const int pl = p > 0 ? p : 1, pb = pl | 1;
    const int MM = (NM - p) > 1 ? (NM - p) : 1;
    if (!P.elig[b]) { if (tid == 0) P.done[b] = 0; return; }
    const double* fs = P.fs + (size_t)b * P.fs_stride;
    const double inv_s = fs[P.fs_stride - 1];
    const int N0 = (int)fs[P.fs_stride - 2], m = (int)fs[P.fs_stride - 3];
    const int N = N0 + m;
    const int base = (p > 0) ? p : N0;
    const double* Ct = fs; const double* M0g = fs + P.off_M0; const double* Gg = fs + P.off_G; const double* Cg = fs + P.off_C;
    const double* Lg = fs + P.off_L;
    // shared: y (N x k), r/u (m x k), s (m x k), L^{-1} packed, G, C
    double* y = smem;                       // N x k   (row i at y + i*k)
    double* rv = y + (size_t)NM * k;        // MM x k
    double* sv = rv + (size_t)MM * k;       // MM x k
    double* t0 = sv + (size_t)MM * k;       // pl x k
    double* Ls = t0 + (size_t)pl * k;       // tri(MM)
    double* Gs = Ls + (size_t)MM * (MM + 1) / 2;
    double* Cs = Gs + (size_t)pb * MM;
    const int* found = P.found + (size_t)b * P.found_stride;
    const int nf = P.n_found[b];
    const int* r4 = P.r4 + (size_t)b * P.r4_stride;
    const double* values = P.values + (size_t)b * P.db_stride * k;
    const double* r3v = P.r3_values ? P.r3_values + (size_t)b * n * k : nullptr;
    for (int e = tid; e < N * k; e += nt) {
        const int i = e / k, q = e % k;
        double val;
        if (i < nf) val = values[(size_t)(found[i] - 1) * k + q];
        else if (i < N0) val = r3v ? r3v[(size_t)(i - nf) * k + q] : 0.0;
        else val = values[(size_t)(r4[i - N0] - 1) * k + q];
        y[e] = val;
    }
    const int tl = (m * (m + 1)) >> 1;
    for (int e = tid; e < tl; e += nt) Ls[e] = Lg[e];
    for (int e = tid; e < pb * m; e += nt) { Gs[e] = Gg[e]; Cs[e] = Cg[e]; }
    double* centers = P.centers + (size_t)b * P.train_stride * n;
    for (int e = tid; e < N * n; e += nt) { const int i = e / n, c = e % n; centers[e] = Ct[(size_t)c * NM + i]; }
    __syncthreads();
    for (int e = tid; e < m * k; e += nt) {                  // r = Y_acc - C' Y_0
        const int eta = e / k, q = e % k;
        double a = y[(size_t)(base + eta) * k + q];
        for (int r = 0; r < p; ++r) a = fma(-Cs[eta * pb + r], y[(size_t)r * k + q], a);
        rv[e] = a;
    }
    __syncthreads();
    for (int e = tid; e < m * k; e += nt) {                  // s = L^{-1} r
        const int r = e / k, q = e % k;
        const double* lr = Ls + ((r * (r + 1)) >> 1);
        double a = 0.0;
        for (int c = 0; c <= r; ++c) a = fma(lr[c], rv[(size_t)c * k + q], a);
        sv[e] = a;
    }
    __syncthreads();
    for (int e = tid; e < m * k; e += nt) {                  // u = L^{-T} s   (into rv)
        const int c = e / k, q = e % k;
        double a = 0.0;
        for (int r = c; r < m; ++r) a = fma(Ls[((r * (r + 1)) >> 1) + c], sv[(size_t)r * k + q], a);
        rv[e] = a;
    }
    __syncthreads();
    double* w_out = P.w + (size_t)b * P.train_stride * k;
    double* lam_out = P.lam + (size_t)b * pl * k;
    for (int e = tid; e < m * k; e += nt) w_out[(size_t)base * k + e] = rv[e];
    for (int e = tid; e < p * k; e += nt) {                  // w_0 = -C u ; t0 = Y_0 - G u
        const int r = e / k, q = e % k;
        double a = 0.0, g = y[(size_t)r * k + q];
        for (int eta = 0; eta < m; ++eta) { const double u = rv[(size_t)eta * k + q]; a = fma(Cs[eta * pb + r], u, a); g = fma(-Gs[eta * pb + r], u, g); }
        w_out[(size_t)r * k + q] = -a;
        t0[e] = g;
    }
    __syncthreads();
    for (int e = tid; e < p * k; e += nt) {                  // lambda~ = M0' t0 -> into sv (p x k)
        const int c = e / k, q = e % k;
        double a = 0.0;
        for (int r = 0; r < p; ++r) a = fma(M0g[r + (size_t)c * pl], t0[(size_t)r * k + q], a);
        sv[e] = a;
    }
    __syncthreads();
    for (int e = tid; e < p * k; e += nt) {                  // back to the monomial basis (1, x_1..x_n)
        const int c = e / k, q = e % k;
        double val;
        if (c == 0) { val = sv[q]; for (int j = 1; j < p; ++j) val = fma(-sv[(size_t)j * k + q] * inv_s, Ct[(size_t)(j - 1) * NM], val); }
        else val = sv[e] * inv_s;
        lam_out[e] = val;
    }
    if (tid == 0) { P.N[b] = N; P.alpha2_out[b] = P.alpha2; P.status[b] = 0; P.done[b] = 1; }
}

// The same build when the kept factorisation is too large for shared memory (large databases: m up to several hundred accepted
// points, packed L^{-1} of ~1 MB per instance): L^{-1} is STREAMED from global memory, twice, row by row with coalesced reads --
// pass 1: s = L^{-1} r (a warp per row, lanes along the row), pass 2: u = L^{-T} s as row axpys into per-warp partial sums.
// Everything else is small and lives in shared memory.  Replaces an O(N^3) from-scratch solve per instance.
__global__ void __launch_bounds__(256) build_prepared_stream_kernel(PreparedBuildParams P) {
    extern __shared__ double smem[];
    const int b = blockIdx.x, n = P.n, k = P.k, p = P.p, NM = P.NM, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int pl = p > 0 ? p : 1, pb = pl | 1;
    const int MM = (NM - p) > 1 ? (NM - p) : 1;
    if (!P.elig[b]) { if (tid == 0) P.done[b] = 0; return; }
    const double* fs = P.fs + (size_t)b * P.fs_stride;
    const double inv_s = fs[P.fs_stride - 1];
    const int N0 = (int)fs[P.fs_stride - 2], m = (int)fs[P.fs_stride - 3];
    const int N = N0 + m;
    const int base = (p > 0) ? p : N0;
    const double* Ct = fs; const double* M0g = fs + P.off_M0; const double* Gg = fs + P.off_G; const double* Cg = fs + P.off_C;
    const double* Lg = fs + P.off_L;
    double* y = smem;                       // NM x k
    double* rv = y + (size_t)NM * k;        // MM x k   r, later u
    double* sv = rv + (size_t)MM * k;       // MM x k   s
    double* t0 = sv + (size_t)MM * k;       // pl x k
    double* up = t0 + (size_t)pl * k;       // nwarps x MM x k partial sums of pass 2
    const int* found = P.found + (size_t)b * P.found_stride;
    const int nf = P.n_found[b];
    const int* r4 = P.r4 + (size_t)b * P.r4_stride;
    const double* values = P.values + (size_t)b * P.db_stride * k;
    const double* r3v = P.r3_values ? P.r3_values + (size_t)b * n * k : nullptr;
    for (int e = tid; e < N * k; e += nt) {
        const int i = e / k, q = e % k;
        double val;
        if (i < nf) val = values[(size_t)(found[i] - 1) * k + q];
        else if (i < N0) val = r3v ? r3v[(size_t)(i - nf) * k + q] : 0.0;
        else val = values[(size_t)(r4[i - N0] - 1) * k + q];
        y[e] = val;
    }
    double* centers = P.centers + (size_t)b * P.train_stride * n;
    for (int e = tid; e < N * n; e += nt) { const int i = e / n, c = e % n; centers[e] = Ct[(size_t)c * NM + i]; }
    for (int e = tid; e < nwarps * m * k; e += nt) up[e] = 0.0;
    __syncthreads();
    for (int e = tid; e < m * k; e += nt) {                  // r = Y_acc - C' Y_0
        const int eta = e / k, q = e % k;
        double a = y[(size_t)(base + eta) * k + q];
        const double* ce = Cg + (size_t)eta * pb;
        for (int r = 0; r < p; ++r) a = fma(-ce[r], y[(size_t)r * k + q], a);
        rv[e] = a;
    }
    __syncthreads();
    for (int r = warp; r < m; r += nwarps) {                 // pass 1: s_r = sum_{c <= r} Linv[r][c] r_c  (k <= 4 outputs at a time)
        const double* lr = Lg + (((size_t)r * (r + 1)) >> 1);
        for (int q0 = 0; q0 < k; q0 += 4) {
            const int kk = min(4, k - q0);
            double a[4] = {0.0, 0.0, 0.0, 0.0};
            for (int c = lane; c <= r; c += 32) {
                const double lv = lr[c];
                for (int q = 0; q < kk; ++q) a[q] = fma(lv, rv[(size_t)c * k + q0 + q], a[q]);
            }
            for (int q = 0; q < kk; ++q) { const double t = warp_sum(a[q]); if (lane == 0) sv[(size_t)r * k + q0 + q] = t; }
        }
    }
    __syncthreads();
    {                                                        // pass 2: u_c = sum_{r >= c} Linv[r][c] s_r, accumulated per warp
        double* mine = up + (size_t)warp * m * k;
        for (int r = warp; r < m; r += nwarps) {
            const double* lr = Lg + (((size_t)r * (r + 1)) >> 1);
            for (int c = lane; c <= r; c += 32) {
                const double lv = lr[c];
                for (int q = 0; q < k; ++q) mine[(size_t)c * k + q] = fma(lv, sv[(size_t)r * k + q], mine[(size_t)c * k + q]);
            }
        }
    }
    __syncthreads();
    for (int e = tid; e < m * k; e += nt) {
        double a = 0.0;
        for (int w = 0; w < nwarps; ++w) a += up[(size_t)w * m * k + e];
        rv[e] = a;
    }
    __syncthreads();
    double* w_out = P.w + (size_t)b * P.train_stride * k;
    double* lam_out = P.lam + (size_t)b * pl * k;
    for (int e = tid; e < m * k; e += nt) w_out[(size_t)base * k + e] = rv[e];
    for (int e = warp; e < p * k; e += nwarps) {             // w_0 = -C u ; t0 = Y_0 - G u   (warp per entry, lanes over the accepted points)
        const int r = e / k, q = e % k;
        double a = 0.0, g = 0.0;
        for (int eta = lane; eta < m; eta += 32) {
            const double u = rv[(size_t)eta * k + q];
            a = fma(Cg[(size_t)eta * pb + r], u, a); g = fma(Gg[(size_t)eta * pb + r], u, g);
        }
        a = warp_sum(a); g = warp_sum(g);
        if (lane == 0) { w_out[(size_t)r * k + q] = -a; t0[e] = y[(size_t)r * k + q] - g; }
    }
    __syncthreads();
    for (int e = tid; e < p * k; e += nt) {                  // lambda~ = M0' t0
        const int c = e / k, q = e % k;
        double a = 0.0;
        for (int r = 0; r < p; ++r) a = fma(M0g[r + (size_t)c * pl], t0[(size_t)r * k + q], a);
        sv[e] = a;
    }
    __syncthreads();
    for (int e = tid; e < p * k; e += nt) {                  // back to the monomial basis (1, x_1..x_n)
        const int c = e / k, q = e % k;
        double val;
        if (c == 0) { val = sv[q]; for (int j = 1; j < p; ++j) val = fma(-sv[(size_t)j * k + q] * inv_s, Ct[(size_t)(j - 1) * NM], val); }
        else val = sv[e] * inv_s;
        lam_out[e] = val;
    }
    if (tid == 0) { P.N[b] = N; P.alpha2_out[b] = P.alpha2; P.status[b] = 0; P.done[b] = 1; }
}

size_t build_prepared_stream_smem_doubles(int n, int k, int NM, int p) {
    (void)n;
    int pl = p > 0 ? p : 1; int MM = (NM - p) > 1 ? (NM - p) : 1;
    return (size_t)NM * k + 2 * (size_t)MM * k + (size_t)pl * k + 8 * (size_t)MM * k + 8;
}
cudaError_t launch_build_prepared_stream(const PreparedBuildParams& P, size_t smem, cudaStream_t s) {
    cudaError_t e = raise_dyn_smem(build_prepared_stream_kernel, smem);
    if (e != cudaSuccess) return e;
    build_prepared_stream_kernel<<<P.B, 256, smem, s>>>(P);
    return cudaGetLastError();
}

size_t build_prepared_smem_doubles(int n, int k, int NM, int p) {
    int pl = p > 0 ? p : 1, pb = pl | 1; int MM = (NM - p) > 1 ? (NM - p) : 1;
    return (size_t)NM * k + 2 * (size_t)MM * k + (size_t)pl * k + (size_t)MM * (MM + 1) / 2 + 2 * (size_t)pb * MM;
}
cudaError_t launch_build_prepared(const PreparedBuildParams& P, size_t smem, cudaStream_t s) {
    cudaError_t e = raise_dyn_smem(build_prepared_kernel, smem);
    if (e != cudaSuccess) return e;
    build_prepared_kernel<<<P.B, 256, smem, s>>>(P);
    return cudaGetLastError();
}


// ------------------------------------------------------------------------------------------------
// Build from the factorisation kept by round4_elim_kernel (mrbf_round4_schur.cu).
//
// That kernel leaves, per instance: M0 = Pi_0^{-T}, the panels C (Lagrange coefficients) and U = Phi(S0, .) - Phi00 C over
// the candidate positions, and the Cholesky factor L of A = N' Phi N over the accepted candidates (column q = q-th accepted
// pivot, rows indexed by candidate position).  With the training order [S0; accepted] (_collect_indices, RbfModel.jl:178-186):
//     r = Y_acc - C_acc' Y_0,   L L' u = r,   w = [-C_acc u; u],   lambda~ = Pi_0^{-1} (Y_0 - U_acc u)
// i.e. two triangular solves per output instead of the O(N^3) saddle-point solve of RBFInterpolationModel (RbfModel.jl:759).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64, 16) build_schur_kernel(SchurBuildParams P) {
    // Two warps per instance and ~6 KB of shared memory: the triangular sweeps are serial chains, so the SM is filled with many
    // instances instead of wide CTAs.  L stays in global memory (L2): the forward sweep walks its columns (contiguous), the
    // backward sweep its rows (one entry per earlier column); both prefetch the next step's entries into registers.
    extern __shared__ double smem[];
    const int b = blockIdx.x, n = P.n, k = P.k, p = P.p, MC = P.MC, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    if (!P.elig[b]) { if (tid == 0) P.done[b] = 0; return; }
    const double* fs = P.fs + (size_t)b * P.fs_stride;
    const double* meta = fs + P.off_acc + MC;
    const double inv_s = meta[0];
    const int N0 = (int)meta[1], m = (int)meta[2], mc = (int)meta[3];
    const int N = N0 + m;
    const double* M0g = fs + P.off_M0; const double* Ug = fs + P.off_U; const double* Cg = fs + P.off_C; const double* Lg = fs + P.off_L;
    double* y0 = smem;                          // p x k
    double* rv = y0 + (size_t)p * k;            // MC x k  (output-major, indexed by candidate position: rv[o * MC + pos])
    double* t0 = rv + (size_t)MC * k;           // p x k
    double* sv = t0 + (size_t)p * k;            // p x k
    double* invd = sv + (size_t)p * k;          // MC  1 / L_qq by accepted ordinal
    int* accpos = reinterpret_cast<int*>(invd + MC);     // ordinal -> position
    int* posq = accpos + MC;                              // position -> ordinal, -1 for rejected candidates
    const int* found = P.found + (size_t)b * P.found_stride;
    const int nf = P.n_found[b];
    const int* r4 = P.r4 + (size_t)b * P.r4_stride;
    const double* sites = P.sites + (size_t)b * P.db_stride * n;
    const double* values = P.values + (size_t)b * P.db_stride * k;
    const double* r3s = P.r3_sites ? P.r3_sites + (size_t)b * n * n : nullptr;
    const double* r3v = P.r3_values ? P.r3_values + (size_t)b * n * k : nullptr;
    for (int i = tid; i < mc; i += nt) posq[i] = -1;
    for (int e = tid; e < mc * k; e += nt) rv[(e / mc) * MC + (e % mc)] = 0.0;
    __syncthreads();
    for (int q = tid; q < m; q += nt) {
        const int j = (int)fs[P.off_acc + q];
        accpos[q] = j; posq[j] = q;
        invd[q] = 1.0 / Lg[(size_t)q * MC + j];
    }
    for (int e = tid; e < p * k; e += nt) {
        const int i = e / k, o = e % k;
        y0[e] = (i < nf) ? values[(size_t)(found[i] - 1) * k + o] : (r3v ? r3v[(size_t)(i - nf) * k + o] : 0.0);
    }
    double* centers = P.centers + (size_t)b * P.train_stride * n;
    // centres = [found; round-3 sites; accepted candidates]: a warp per row, four rows in flight (id lookup, coalesced row read, store)
    for (int i0 = 4 * warp; i0 < N; i0 += 4 * nwarps) {
        const double* src[4];
#pragma unroll
        for (int u_ = 0; u_ < 4; ++u_) {
            const int i = i0 + u_;
            src[u_] = (i >= N) ? nullptr : ((i < nf) ? sites + (size_t)(found[i] - 1) * n : ((i < N0) ? r3s + (size_t)(i - nf) * n : sites + (size_t)(r4[i - N0] - 1) * n));
        }
        for (int c = lane; c < n; c += 32) {
            double v[4];
#pragma unroll
            for (int u_ = 0; u_ < 4; ++u_) v[u_] = src[u_] ? src[u_][c] : 0.0;
#pragma unroll
            for (int u_ = 0; u_ < 4; ++u_) if (src[u_]) centers[(size_t)(i0 + u_) * n + c] = v[u_];
        }
    }
    __syncthreads();
    for (int e = tid; e < m * k; e += nt) {      // r = Y_acc - C_acc' Y_0, stored by candidate position
        const int q = e % m, o = e / m;
        double a = values[(size_t)(r4[q] - 1) * k + o];
        const double* cq = Cg + accpos[q];
        for (int r = 0; r < p; ++r) a = fma(-cq[(size_t)r * MC], y0[r * k + o], a);
        rv[o * MC + accpos[q]] = a;
    }
    __syncthreads();
    // L t = r, then L' u = t.  One warp per output; the right-hand side stays in registers (lane l owns positions l, l + 32,
    // l + 64, l + 96 -- mc <= 128), the pivot entry travels by one shuffle per step and both sweeps are axpy updates.
    for (int o = warp; o < k; o += nwarps) {
        double* r = rv + o * MC;
        const int i0 = lane, i1 = lane + 32, i2 = lane + 64, i3 = lane + 96;
        double r0 = (i0 < mc) ? r[i0] : 0.0, r1 = (i1 < mc) ? r[i1] : 0.0, r2 = (i2 < mc) ? r[i2] : 0.0, r3 = (i3 < mc) ? r[i3] : 0.0;
        double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;                    // column of the current step, prefetched
        if (m > 0) {                                                           // only the rows below the pivot are ever used
            const double* col = Lg; const int jn = accpos[0];
            c0 = (i0 > jn && i0 < mc) ? col[i0] : 0.0; c1 = (i1 > jn && i1 < mc) ? col[i1] : 0.0;
            c2 = (i2 > jn && i2 < mc) ? col[i2] : 0.0; c3 = (i3 > jn && i3 < mc) ? col[i3] : 0.0;
        }
        for (int q = 0; q < m; ++q) {
            double n0 = 0.0, n1 = 0.0, n2 = 0.0, n3 = 0.0;
            if (q + 1 < m) {                                                   // next column on its way while this one is used
                const double* col = Lg + (size_t)(q + 1) * MC; const int jn = accpos[q + 1];
                n0 = (i0 > jn && i0 < mc) ? col[i0] : 0.0; n1 = (i1 > jn && i1 < mc) ? col[i1] : 0.0;
                n2 = (i2 > jn && i2 < mc) ? col[i2] : 0.0; n3 = (i3 > jn && i3 < mc) ? col[i3] : 0.0;
            }
            const int j = accpos[q], sl = j >> 5;
            const double mine = (sl == 0) ? r0 : ((sl == 1) ? r1 : ((sl == 2) ? r2 : r3));
            const double t = __shfl_sync(0xffffffffu, mine, j & 31) * invd[q];
            if (i0 == j) r0 = t; else if (i0 > j) r0 = fma(-c0, t, r0);
            if (i1 == j) r1 = t; else if (i1 > j) r1 = fma(-c1, t, r1);
            if (i2 == j) r2 = t; else if (i2 > j) r2 = fma(-c2, t, r2);
            if (i3 == j) r3 = t; else if (i3 > j) r3 = fma(-c3, t, r3);
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        }
        // rejected positions carry no unknown.  Backward sweep column by column as well (contiguous, coalesced reads of L; a row-wise
        // axpy would touch one 32-byte sector per entry): u_j = (t_j - sum_{i > j} L[i][j] u_i) / L_jj with a warp reduction per step.
        if (!(i0 < mc && posq[i0] >= 0)) r0 = 0.0;
        if (!(i1 < mc && posq[i1] >= 0)) r1 = 0.0;
        if (!(i2 < mc && posq[i2] >= 0)) r2 = 0.0;
        if (!(i3 < mc && posq[i3] >= 0)) r3 = 0.0;
        if (m > 0) {
            const double* col = Lg + (size_t)(m - 1) * MC; const int jn = accpos[m - 1];
            c0 = (i0 > jn && i0 < mc) ? col[i0] : 0.0; c1 = (i1 > jn && i1 < mc) ? col[i1] : 0.0;
            c2 = (i2 > jn && i2 < mc) ? col[i2] : 0.0; c3 = (i3 > jn && i3 < mc) ? col[i3] : 0.0;
        }
        for (int q = m - 1; q >= 0; --q) {
            double n0 = 0.0, n1 = 0.0, n2 = 0.0, n3 = 0.0;
            if (q > 0) {
                const double* col = Lg + (size_t)(q - 1) * MC; const int jn = accpos[q - 1];
                n0 = (i0 > jn && i0 < mc) ? col[i0] : 0.0; n1 = (i1 > jn && i1 < mc) ? col[i1] : 0.0;
                n2 = (i2 > jn && i2 < mc) ? col[i2] : 0.0; n3 = (i3 > jn && i3 < mc) ? col[i3] : 0.0;
            }
            const int j = accpos[q];
            double part = 0.0;
            if (i0 > j) part = fma(c0, r0, part);
            if (i1 > j) part = fma(c1, r1, part);
            if (i2 > j) part = fma(c2, r2, part);
            if (i3 > j) part = fma(c3, r3, part);
            part = warp_sum(part);
            const double id_ = invd[q];
            if (i0 == j) r0 = (r0 - part) * id_;
            if (i1 == j) r1 = (r1 - part) * id_;
            if (i2 == j) r2 = (r2 - part) * id_;
            if (i3 == j) r3 = (r3 - part) * id_;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        }
        if (i0 < mc) r[i0] = r0;
        if (i1 < mc) r[i1] = r1;
        if (i2 < mc) r[i2] = r2;
        if (i3 < mc) r[i3] = r3;
    }
    __syncthreads();
    double* w_out = P.w + (size_t)b * P.train_stride * k;
    double* lam_out = P.lam + (size_t)b * p * k;
    for (int e = tid; e < m * k; e += nt) { const int q = e / k, o = e % k; w_out[(size_t)(p + q) * k + o] = rv[o * MC + accpos[q]]; }
    // w_0 = -C_acc u ; t0 = Y_0 - U_acc u : a warp takes four rows of C and U at a time (lanes over the accepted points), so eight row
    // reads are in flight and the eight reductions overlap -- one (row, output) entry per pass was a chain of DRAM latencies
    const int nrc = (p + 3) >> 2;
    for (int e = warp; e < nrc * k; e += nwarps) {
        const int r0 = 4 * (e / k), o = e % k;
        const double* u = rv + o * MC;
        double a[4] = {0.0, 0.0, 0.0, 0.0}, g_[4] = {0.0, 0.0, 0.0, 0.0};
        for (int q = lane; q < m; q += 32) {
            const int pos = accpos[q];
            const double uv = u[pos];
#pragma unroll
            for (int u_ = 0; u_ < 4; ++u_) {
                const int r = min(r0 + u_, p - 1);
                a[u_] = fma(Cg[(size_t)r * MC + pos], uv, a[u_]); g_[u_] = fma(Ug[(size_t)r * MC + pos], uv, g_[u_]);
            }
        }
#pragma unroll
        for (int u_ = 0; u_ < 4; ++u_) { a[u_] = warp_sum(a[u_]); g_[u_] = warp_sum(g_[u_]); }
        if (lane == 0) {
#pragma unroll
            for (int u_ = 0; u_ < 4; ++u_) {
                const int r = r0 + u_;
                if (r < p) { w_out[(size_t)r * k + o] = -a[u_]; t0[r * k + o] = y0[r * k + o] - g_[u_]; }
            }
        }
    }
    __syncthreads();
    for (int e = tid; e < p * k; e += nt) {      // lambda~ = Pi_0^{-1} t0 = M0' t0
        const int c = e / k, o = e % k;
        double a = 0.0;
        for (int r = 0; r < p; ++r) a = fma(M0g[r + (size_t)c * p], t0[r * k + o], a);
        sv[e] = a;
    }
    __syncthreads();
    for (int e = tid; e < p * k; e += nt) {      // back from the centred / scaled basis to the monomials (1, x_1..x_n)
        const int c = e / k, o = e % k;
        double val;
        if (c == 0) { val = sv[o]; for (int j = 1; j < p; ++j) val = fma(-sv[(size_t)j * k + o] * inv_s, centers[j - 1], val); }
        else val = sv[e] * inv_s;
        lam_out[e] = val;
    }
    if (tid == 0) { P.N[b] = N; if (P.alpha2 >= 0.0) P.alpha2_out[b] = P.alpha2; P.status[b] = 0; P.done[b] = 1; }
}

size_t build_schur_smem_doubles(int k, int MC, int p) {
    return 3 * (size_t)p * k + (size_t)MC * k + MC + MC + 2;
}
cudaError_t launch_build_schur(const SchurBuildParams& P, size_t smem, cudaStream_t s) {
    cudaError_t e = raise_dyn_smem(build_schur_kernel, smem);
    if (e != cudaSuccess) return e;
    build_schur_kernel<<<P.B, 64, smem, s>>>(P);
    return cudaGetLastError();
}

size_t build_vec_doubles(int n, int k, int ld, int p) { int pl = p > 0 ? p : 1; return 2 * (size_t)ld + pl + 80; }
size_t build_ws_doubles(int n, int k, int ld, int p) { int pl = p > 0 ? p : 1; return (size_t)ld * ld + (size_t)ld * pl + (size_t)ld * k; }

template <int NT>
static cudaError_t launch_build_nt(const BuildParams& P, size_t smem, cudaStream_t s) {
    cudaError_t e = raise_dyn_smem(build_kernel<NT>, smem);
    if (e != cudaSuccess) return e;
    build_kernel<NT><<<P.B, NT, smem, s>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_build(const BuildParams& P, size_t smem, cudaStream_t s) {
    const bool few_large = P.B <= 148 && !P.ws_in_smem;
    int nt = few_large ? 1024 : 256;
    // a system that needs more than half of the SM's shared memory runs one CTA per SM anyway: more threads per CTA then shorten
    // every barrier phase at no cost in occupancy
    if (!few_large && smem > 113 * 1024) nt = 1024;
    if (const char* ev = getenv("MRBF_BUILD_NT")) { const int v = atoi(ev); if (v == 256 || v == 512 || v == 1024) nt = v; }
    return nt == 1024 ? launch_build_nt<1024>(P, smem, s) : (nt == 512 ? launch_build_nt<512>(P, smem, s) : launch_build_nt<256>(P, smem, s));
}

}  // namespace mrbf
