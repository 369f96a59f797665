// Round 4 (Wild's bounded-Cholesky augmentation, RbfModel.jl:352-499) as a right-looking elimination held in registers.
//
// Same decisions as the reference loop (RbfModel.jl:420-452), restated -- not copied -- for a CTA per instance:
//
//   * regular case only: the found set S0 = [centre; r1; r2; r3] has exactly p = n + 1 points and Pi_0 = Pi(S0) is
//     non-singular.  Every candidate xi then has the null vector n_xi = e_xi - sum_s c_xi[s] e_s with the Lagrange
//     coefficients c_xi = Pi_0^{-T} pi_xi, and the vectors n_xi of the accepted points span null(Pi') like the
//     reference's orthonormal Z (RbfModel.jl:391, 465).
//   * the reference tests   tau^2 = sigma - ||L^{-1} v||^2 > theta^4   (RbfModel.jl:447-452).  In the basis above
//         tau^2 = d^2 / (1 + lev),
//     d^2  = Schur complement of  A = N' Phi N  at xi after eliminating the accepted candidates,
//     1+lev = Schur complement of W = I + C' C   at xi after eliminating the accepted candidates
//     (lev = pi' (Pi' Pi)^{-1} pi is the leverage whose 1/(1+lev) is the product of the Givens cosines of
//     utilities.jl:437-448; Woodbury turns the Sherman-Morrison updates of (Pi' Pi)^{-1} into an elimination on W).
//   * so round 4 is ONE symmetric elimination on the pair (A, W) over the candidates in ascending id order, where a
//     rejected pivot is simply skipped (RbfModel.jl:452 leaves the state untouched).  A and W (mc x mc, mc <= 128
//     candidates) never touch memory: every thread owns one 4 x 4 tile of each in registers, computed directly from
//     the shared-memory panels C (Lagrange coefficients), V = B - Phi00 C / 2 and the candidate sites
//         A_ij = phi(|xi_i - xi_j|) - c_i.v_j - v_i.c_j ,      W_ij = delta_ij + c_i.c_j .
//     Per pivot: the tile column owners publish the pivot column (64 B per thread), one barrier, every live tile does
//     a 4 x 4 rank-1 update of A and of W.  The Cholesky factor column (pivot column / d) is streamed out for
//     mrbf_build_prepared_dev, which finishes the model with two triangular solves per output.
//
// Instances that do not qualify (N0 != p, singular Pi_0) are marked n_r4 = -1 for the literal kernel; batches whose
// database is larger than 128 sites or whose panels do not fit in shared memory use round4_block_kernel instead.
#include "mrbf_common.cuh"
#include "mrbf_kernels.h"

namespace mrbf {

__host__ __device__ inline int schur_tiles(int TR) { return (TR * (TR + 1)) / 2; }

SchurGeom round4_schur_geom(int n, int p, int db_stride) {
    SchurGeom g{};
    const int pl = p > 0 ? p : 1;
    g.MC = (db_stride + 3) & ~3;
    if (g.MC < 4) g.MC = 4;
    g.TR = g.MC / 4;
    g.ntiles = schur_tiles(g.TR);
    int nt = (g.ntiles + 31) & ~31;
    if (nt < 256) nt = 256;
    g.nthreads = nt;
    auto up4 = [](size_t v) { return (v + 3) & ~(size_t)3; };          // every block starts 32-byte aligned (double4 accesses)
    const size_t cv = up4((size_t)2 * pl * (g.MC > pl ? g.MC : pl));   // [C | V] panels; Gauss-Jordan scratch aliases them
    g.sm_C = 0; g.sm_V = (size_t)pl * g.MC;
    g.sm_Xc = cv; g.sm_X0 = g.sm_Xc + (size_t)n * g.MC;
    g.sm_M0 = up4(g.sm_X0 + (size_t)pl * n); g.sm_P00 = up4(g.sm_M0 + (size_t)pl * pl);
    g.sm_col = up4(g.sm_P00 + (size_t)pl * pl);                         // colA[2][MC], colW[2][MC], pivot values[2][4]
    g.sm_red = g.sm_col + (size_t)4 * g.MC + 8;
    g.sm_int = g.sm_red + 80;                                           // clist[MC] ints, wcnt[32] ints, flags[db_stride] bytes
    const size_t ints = (size_t)g.MC + 32 + ((size_t)db_stride + 3) / 4 + 4;
    g.smem_doubles = g.sm_int + (ints + 1) / 2;
    // kept state per instance: M0 (p x p), U (p x MC), C (p x MC), L (MC x MC, column q = q-th accepted pivot), accpos (MC), meta (8)
    g.off_M0 = 0; g.off_U = up4((size_t)pl * pl); g.off_C = g.off_U + (size_t)pl * g.MC; g.off_L = g.off_C + (size_t)pl * g.MC;
    g.off_acc = g.off_L + (size_t)g.MC * g.MC; g.state_doubles = up4(g.off_acc + g.MC + 8);
    g.eligible = (p > 0 && g.MC <= 128 && g.nthreads <= 544 && g.smem_doubles * sizeof(double) <= (size_t)225 * 1024) ? 1 : 0;
    return g;
}

__global__ void __launch_bounds__(544, 1) round4_schur_kernel(Round4Params P, SchurGeom g) {
    extern __shared__ double smem[];
    const int b = blockIdx.x, n = P.n, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int p = poly_dim(n, P.cfg.polynomial_degree), pl = p;
    const int MC = g.MC, TR = g.TR, MQ = MC >> 2;
    double* Cs = smem + g.sm_C; double* Vs = smem + g.sm_V; double* Xc = smem + g.sm_Xc; double* X0 = smem + g.sm_X0;
    double* M0 = smem + g.sm_M0; double* P00 = smem + g.sm_P00;
    double* colA = smem + g.sm_col; double* colW = colA + 2 * MC; double* diag = colW + 2 * MC;
    double* red = smem + g.sm_red;
    int* clist = reinterpret_cast<int*>(smem + g.sm_int); int* wcnt = clist + MC;
    unsigned char* cflag = reinterpret_cast<unsigned char*>(wcnt + 32);
    double* Aq = smem;                 // Gauss-Jordan scratch [Pi_0 | I] (p x 2p), dead before C is written
    double* Qx = Aq + pl * pl;

    const int n_db = P.n_db[b];
    const double* sites = P.sites + (size_t)b * P.db_stride * n;
    const double* lb2 = P.lb2 + (size_t)b * n;
    const double* ub2 = P.ub2 + (size_t)b * n;
    const int* found = P.found + (size_t)b * P.found_stride;
    const int nf_ids = P.n_found[b];
    const int n_extra = P.n_extra ? P.n_extra[b] : 0;
    const double* extra = P.extra_sites ? P.extra_sites + (size_t)b * P.extra_stride * n : nullptr;
    int* r4 = P.r4 + (size_t)b * P.r4_stride;
    const int N0 = nf_ids + n_extra;
    const int max_points = P.max_points;
    if (tid == 0 && P.elig) P.elig[b] = 0;
    if (!(N0 < max_points)) { if (tid == 0) { P.n_r4[b] = 0; if (P.status) P.status[b] = 0; } return; }
    if (N0 != p || n_db > MC) { if (tid == 0) P.n_r4[b] = -1; return; }      // literal kernel takes over

    // ---- candidates: results_in_box_indices(db, lb_2, ub_2, found) in ascending id order (RbfModel.jl:360)
    for (int id = tid; id < n_db; id += nt) {
        bool ok = in_box_pt(sites + (size_t)id * n, lb2, ub2, n);
        for (int f = 0; f < nf_ids && ok; ++f) ok = (found[f] != id + 1);
        cflag[id] = ok ? 1 : 0;
    }
    for (int e = tid; e < p * n; e += nt) {
        const int i = e / n, k = e % n;
        X0[e] = (i < nf_ids) ? sites[(size_t)(found[i] - 1) * n + k] : extra[(size_t)(i - nf_ids) * n + k];
    }
    if (tid == 0) red[76] = 0.0;
    __syncthreads();
    int mc = 0;
    {
        const int nseg = (n_db + 31) >> 5;         // <= 4 segments of 32 ids
        if (warp < nseg) {
            const int id = warp * 32 + lane;
            const bool f = id < n_db && cflag[id];
            const unsigned msk = __ballot_sync(0xffffffffu, f);
            if (lane == 0) wcnt[warp] = __popc(msk);
        }
        __syncthreads();
        int basew = 0;
        for (int w = 0; w < nseg; ++w) { if (w < warp) basew += wcnt[w]; mc += wcnt[w]; }
        if (warp < nseg) {
            const int id = warp * 32 + lane;
            const bool f = id < n_db && cflag[id];
            const unsigned msk = __ballot_sync(0xffffffffu, f);
            if (f) clist[basew + __popc(msk & ((1u << lane) - 1u))] = id;
        }
    }
    // The polynomial basis is centred at the first found point and scaled by the spread of S0: c_xi and the leverage are
    // invariant under that change of basis, and Pi_0 stays well conditioned for tiny Delta.
    double inv_s = 1.0;
    if (p > 1) {
        double mx = 0.0;
        for (int e = tid; e < p * n; e += nt) { const int k = e % n; mx = fmax(mx, fabs(X0[e] - X0[k])); }
        mx = warp_max(mx);
        if (lane == 0) red[40 + warp] = mx;
        __syncthreads();
        mx = 0.0;
        for (int w = 0; w < nwarps; ++w) mx = fmax(mx, red[40 + w]);
        inv_s = mx > 0.0 ? 1.0 / mx : 1.0;
    }
    __syncthreads();
    if (mc == 0) { if (tid == 0) { P.n_r4[b] = 0; if (P.status) P.status[b] = 0; } return; }
    // candidate sites, coordinate-major (Xc[k][i]); the padding columns repeat the centre (finite, never a pivot)
    for (int e = tid; e < MC * n; e += nt) {
        const int i = e / n, k = e % n;
        Xc[k * MC + i] = (i < mc) ? sites[(size_t)clist[i] * n + k] : X0[k];
    }
    for (int e = tid; e < p * p; e += nt) {
        const int i = e % p, j = e / p;
        double r2 = 0.0;
        for (int k = 0; k < n; ++k) { const double d = X0[i * n + k] - X0[j * n + k]; r2 = fma(d, d, r2); }
        P00[i + j * pl] = rad_phi(P.rf, r2);
        Aq[i + j * pl] = (j == 0) ? 1.0 : (X0[i * n + j - 1] - X0[j - 1]) * inv_s;
        Qx[i + j * pl] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    // ---- Pi_0^{-1} by Gauss-Jordan with partial pivoting on [Pi_0 | I]
    for (int kk = 0; kk < p; ++kk) {
        if (warp == 0) {
            ArgMax mine; mine.v = 0.0; mine.id = -1;
            for (int i = kk + lane; i < p; i += 32) { ArgMax c_; c_.v = fabs(Aq[i + kk * pl]); c_.id = i; mine = better(mine, c_); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ArgMax t_; t_.v = __shfl_xor_sync(0xffffffffu, mine.v, o); t_.id = __shfl_xor_sync(0xffffffffu, mine.id, o);
                mine = better(mine, t_);
            }
            if (!(mine.v > 1e-12)) { if (lane == 0) red[76] = 1.0; }       // Pi_0 (scaled to O(1)) is rank deficient
            else {
                const double rp = 1.0 / Aq[mine.id + kk * pl];
                __syncwarp();
                for (int c = kk + lane; c < 2 * p; c += 32) {
                    const double a = Aq[mine.id + c * pl], bq = Aq[kk + c * pl];
                    Aq[mine.id + c * pl] = bq; Aq[kk + c * pl] = a * rp;
                }
            }
        }
        __syncthreads();
        if (red[76] != 0.0) { if (tid == 0) P.n_r4[b] = -1; return; }
        for (int c = kk + 1 + warp; c < 2 * p; c += nwarps) {
            const double pk = Aq[kk + c * pl];
            for (int i = lane; i < p; i += 32) if (i != kk) Aq[i + c * pl] = fma(-Aq[i + kk * pl], pk, Aq[i + c * pl]);
        }
        __syncthreads();
    }
    for (int e = tid; e < p * p; e += nt) { const int r = e % p, c = e / p; M0[r + c * pl] = Qx[c + r * pl]; }   // M0 = Pi_0^{-T}
    __syncthreads();                                // Gauss-Jordan scratch is dead from here on
    double* keep = P.keep_fs ? P.keep_fs + (size_t)b * P.fs_stride : nullptr;
    if (keep) for (int e = tid; e < p * p; e += nt) keep[g.off_M0 + e] = M0[e];

    // ---- panels: C = Pi_0^{-T} pi~ (Lagrange coefficients) and B = Phi(S0, candidates), one (row, 4 candidates) task per thread
    for (int t = tid; t < p * MQ; t += nt) {
        const int r = t / MQ, i4 = (t % MQ) * 4;
        double c0 = M0[r], c1 = c0, c2 = c0, c3 = c0;            // pi~[0] = 1
        double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
        for (int k = 0; k < n; ++k) {
            const double4 x = *reinterpret_cast<const double4*>(Xc + k * MC + i4);
            if (k + 1 < p) {
                const double mv = M0[r + (k + 1) * pl] * inv_s, xc = X0[k];
                c0 = fma(mv, x.x - xc, c0); c1 = fma(mv, x.y - xc, c1); c2 = fma(mv, x.z - xc, c2); c3 = fma(mv, x.w - xc, c3);
            }
            const double xr = X0[r * n + k];
            double e_;
            e_ = x.x - xr; d0 = fma(e_, e_, d0); e_ = x.y - xr; d1 = fma(e_, e_, d1);
            e_ = x.z - xr; d2 = fma(e_, e_, d2); e_ = x.w - xr; d3 = fma(e_, e_, d3);
        }
        *reinterpret_cast<double4*>(Cs + r * MC + i4) = make_double4(c0, c1, c2, c3);
        *reinterpret_cast<double4*>(Vs + r * MC + i4) = make_double4(rad_phi(P.rf, d0), rad_phi(P.rf, d1), rad_phi(P.rf, d2), rad_phi(P.rf, d3));
        if (keep) *reinterpret_cast<double4*>(keep + g.off_C + r * MC + i4) = make_double4(c0, c1, c2, c3);
    }
    __syncthreads();
    // ---- U = B - Phi00 C (kept for the build), V = B - Phi00 C / 2 (in place of B)
    for (int t = tid; t < p * MQ; t += nt) {
        const int r = t / MQ, i4 = (t % MQ) * 4;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        for (int c = 0; c < p; ++c) {
            const double pv = P00[r + c * pl];
            const double4 cc = *reinterpret_cast<const double4*>(Cs + c * MC + i4);
            s0 = fma(pv, cc.x, s0); s1 = fma(pv, cc.y, s1); s2 = fma(pv, cc.z, s2); s3 = fma(pv, cc.w, s3);
        }
        const double4 bb = *reinterpret_cast<const double4*>(Vs + r * MC + i4);
        if (keep) *reinterpret_cast<double4*>(keep + g.off_U + r * MC + i4) = make_double4(bb.x - s0, bb.y - s1, bb.z - s2, bb.w - s3);
        *reinterpret_cast<double4*>(Vs + r * MC + i4) = make_double4(fma(-0.5, s0, bb.x), fma(-0.5, s1, bb.y), fma(-0.5, s2, bb.z), fma(-0.5, s3, bb.w));
    }
    __syncthreads();

    // ---- tiles: thread t owns tile (I, K), I >= K, of A and of W.  Tiles are numbered column by column from the LAST tile
    // column, so the tiles that are still live at pivot j (K >= j / 4) are always a prefix of the thread block.
    int tI = 0, tK = 0;
    const bool has_tile = tid < g.ntiles;
    if (has_tile) {
        int s = 0;
        while (((s + 1) * (s + 2)) / 2 <= tid) ++s;
        tK = TR - 1 - s; tI = tK + (tid - (s * (s + 1)) / 2);
    }
    double A[4][4], W[4][4];
    if (has_tile) {
        const double* xi = Xc + 4 * tI; const double* xk = Xc + 4 * tK;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) A[a][c] = 0.0;
        for (int k = 0; k < n; ++k) {
            const double4 vi = *reinterpret_cast<const double4*>(xi + k * MC);
            const double4 vk = *reinterpret_cast<const double4*>(xk + k * MC);
            const double ri[4] = {vi.x, vi.y, vi.z, vi.w}, rk[4] = {vk.x, vk.y, vk.z, vk.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) { const double d = ri[a] - rk[c]; A[a][c] = fma(d, d, A[a][c]); }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) { A[a][c] = rad_phi(P.rf, A[a][c]); W[a][c] = (tI == tK && a == c) ? 1.0 : 0.0; }
        const double* ci_ = Cs + 4 * tI; const double* ck_ = Cs + 4 * tK; const double* vi_ = Vs + 4 * tI; const double* vk_ = Vs + 4 * tK;
        for (int r = 0; r < p; ++r) {
            const double4 a4 = *reinterpret_cast<const double4*>(ci_ + r * MC), b4 = *reinterpret_cast<const double4*>(ck_ + r * MC);
            const double4 c4 = *reinterpret_cast<const double4*>(vi_ + r * MC), d4 = *reinterpret_cast<const double4*>(vk_ + r * MC);
            const double ci[4] = {a4.x, a4.y, a4.z, a4.w}, ck[4] = {b4.x, b4.y, b4.z, b4.w};
            const double vi[4] = {c4.x, c4.y, c4.z, c4.w}, vk[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    A[a][c] = fma(-ci[a], vk[c], A[a][c]);
                    A[a][c] = fma(-vi[a], ck[c], A[a][c]);
                    W[a][c] = fma(ci[a], ck[c], W[a][c]);
                }
        }
    }
    // Publishing pivot column jn: the owners of tile column jn / 4 write their entries below the diagonal (zeros above, so
    // the rank-1 update needs no masks); the owner of the diagonal entry also publishes d^2, 1 + lev and their reciprocal
    // roots -- the only divisions of the step, done by ONE thread instead of by the whole block.
    //   pv[0] = d^2, pv[1] = 1 + lev, pv[2] = 1 / d, pv[3] = 1 / (1 + lev)
    double* pvs = diag;                             // 2 x 4 doubles
#define SCHUR_PUBLISH(SLOT, JN, CN)                                                                              \
    {                                                                                                            \
        double* na_ = colA + (SLOT) * MC; double* nw_ = colW + (SLOT) * MC;                                      \
        _Pragma("unroll") for (int a = 0; a < 4; ++a) {                                                          \
            const int i = 4 * tI + a;                                                                            \
            na_[i] = (i > (JN)) ? A[a][CN] : 0.0; nw_[i] = (i > (JN)) ? W[a][CN] : 0.0;                           \
            if (i == (JN)) {                                                                                     \
                const double da_ = A[a][CN], dw_ = W[a][CN];                                                     \
                pvs[4 * (SLOT) + 0] = da_; pvs[4 * (SLOT) + 1] = dw_;                                            \
                pvs[4 * (SLOT) + 2] = rsqrt(da_); pvs[4 * (SLOT) + 3] = 1.0 / dw_;                                \
            }                                                                                                    \
        }                                                                                                        \
    }
    if (has_tile && tK == 0) SCHUR_PUBLISH(0, 0, 0)
    __syncthreads();

    // ---- elimination over the candidates in ascending id order
    const double thr = P.chol_thr;
    int nacc = 0;
    bool full = false;
    for (int j0 = 0; j0 < mc && !full; j0 += 4) {
        const int Kj = j0 >> 2;
        const int live = schur_tiles(TR - Kj);       // tiles with K >= Kj
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = j0 + jj;
            if (j >= mc || full) break;
            const int cur = jj & 1, nxt = cur ^ 1;
            const double dA = pvs[4 * cur], rw = pvs[4 * cur + 3];
            const double tau2 = dA * rw;            // d^2 / (1 + lev) == sigma - ||L^-1 v||^2 of RbfModel.jl:447-449
            if (tau2 > thr) {                        // RbfModel.jl:452
                const double* ca = colA + cur * MC; const double* cw = colW + cur * MC;
                const double rd = pvs[4 * cur + 2];
                if (tid < live) {
                    const double ra = rd * rd;
                    {
                        const double4 i4 = *reinterpret_cast<const double4*>(ca + 4 * tI), k4 = *reinterpret_cast<const double4*>(ca + 4 * tK);
                        const double ai[4] = {i4.x * ra, i4.y * ra, i4.z * ra, i4.w * ra}, ak[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                        for (int a = 0; a < 4; ++a)
#pragma unroll
                            for (int c = 0; c < 4; ++c) A[a][c] = fma(-ai[a], ak[c], A[a][c]);
                    }
                    {
                        const double4 i4 = *reinterpret_cast<const double4*>(cw + 4 * tI), k4 = *reinterpret_cast<const double4*>(cw + 4 * tK);
                        const double wi[4] = {i4.x * rw, i4.y * rw, i4.z * rw, i4.w * rw}, wk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                        for (int a = 0; a < 4; ++a)
#pragma unroll
                            for (int c = 0; c < 4; ++c) W[a][c] = fma(-wi[a], wk[c], W[a][c]);
                    }
                }
                if (keep) {                          // column nacc of the Cholesky factor of A: pivot column / d, d on the diagonal
                    double* Lc = keep + g.off_L + (size_t)nacc * MC;
                    for (int i = j + tid; i < mc; i += nt) Lc[i] = (i == j) ? dA * rd : ca[i] * rd;
                    if (tid == 0) keep[g.off_acc + nacc] = (double)j;
                }
                if (tid == 0) r4[nacc] = clist[j] + 1;
                nacc += 1;
                full = !(N0 + nacc < max_points) || !(nacc < P.r4_stride);       // RbfModel.jl:402
            }
            // publish pivot column j + 1 (final after this update): owners are the tiles of tile column (j + 1) / 4
            if (j + 1 < mc && !full) {
                const int cn = (jj + 1) & 3, Kn = (jj == 3) ? Kj + 1 : Kj;
                if (has_tile && tK == Kn) SCHUR_PUBLISH(nxt, j + 1, cn)
            }
            __syncthreads();
        }
    }
#undef SCHUR_PUBLISH
    if (tid == 0) {
        P.n_r4[b] = nacc; if (P.status) P.status[b] = 0;
        if (keep) {
            keep[g.off_acc + MC + 0] = inv_s; keep[g.off_acc + MC + 1] = (double)N0; keep[g.off_acc + MC + 2] = (double)nacc;
            keep[g.off_acc + MC + 3] = (double)mc;
            P.elig[b] = 1;
        }
    }
}

cudaError_t launch_round4_schur(const Round4Params& P, const SchurGeom& g, cudaStream_t s) {
    const size_t smem = g.smem_doubles * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(round4_schur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    round4_schur_kernel<<<P.B, g.nthreads, smem, s>>>(P, g);
    return cudaGetLastError();
}

}  // namespace mrbf
