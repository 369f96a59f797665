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
//     The elimination is blocked, four pivots (one tile column) at a time: the diagonal-tile thread runs the four pivot
//     tests in registers, the tiles of that tile column publish the pivot panel, every tile to the right applies the
//     rank-4 update; the hand-overs are split-phase mbarriers.  The Cholesky factor rows (pivot column / d) are streamed
//     out for mrbf_build_prepared_dev, which finishes the model with two triangular solves per output.
//   * two kernels: round4_panels_kernel (candidate list, Pi_0^{-1}, panels; four instances per SM) hands the panels over
//     through an L2-resident workspace to round4_schur_kernel (tiles + elimination; the tiles fill the register file, one
//     instance per SM, in two launch shapes: <= 100 candidates on 352 threads / 157 registers, else 544 threads).
//
// Instances that do not qualify (N0 != p, singular Pi_0) are marked n_r4 = -1 for the literal kernel; batches whose
// database is larger than 128 sites or whose panels do not fit in shared memory use round4_block_kernel instead.
#include "mrbf_common.cuh"
#include "mrbf_kernels.h"

namespace mrbf {

__host__ __device__ inline int schur_tiles(int TR) { return (TR * (TR + 1)) / 2; }

// split-phase barrier (mbarrier in shared memory): arrive does not block, wait spins on the phase parity
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_addr(bar)) : "memory");
}
// Row index (lo <= i < hi) of the entry of largest magnitude in col[], -1 if none exceeds 1e-12.  One warp; the magnitudes
// are compared as integers (IEEE doubles order like their bit patterns) with redux.sync when the range fits a warp.
__device__ __forceinline__ int warp_argmax_abs(const double* col, int lo, int hi, int lane) {
    if (hi - lo <= 32) {
        const int i = lo + lane;
        const double v = (i < hi) ? fabs(col[i]) : 0.0;
        const unsigned vh = (unsigned)__double2hiint(v), vl = (unsigned)__double2loint(v);
        const unsigned mh = __reduce_max_sync(0xffffffffu, vh);
        const unsigned ml = __reduce_max_sync(0xffffffffu, vh == mh ? vl : 0u);
        const unsigned win = __ballot_sync(0xffffffffu, vh == mh && vl == ml);
        const double mv = __hiloint2double((int)mh, (int)ml);
        return (mv > 1e-12) ? lo + (__ffs(win) - 1) : -1;
    }
    ArgMax mine; mine.v = 0.0; mine.id = -1;
    for (int i = lo + lane; i < hi; i += 32) { ArgMax c_; c_.v = fabs(col[i]); c_.id = i; mine = better(mine, c_); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgMax t_; t_.v = __shfl_xor_sync(0xffffffffu, mine.v, o); t_.id = __shfl_xor_sync(0xffffffffu, mine.id, o);
        mine = better(mine, t_);
    }
    return (mine.v > 1e-12) ? mine.id : -1;
}
__device__ __forceinline__ void mbar_arrive_drop(unsigned long long* bar) {
    asm volatile("mbarrier.arrive_drop.shared::cta.b64 _, [%0];" :: "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    } while (!done);
}

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
    const size_t ints = (size_t)g.MC + 32 + ((size_t)db_stride + 3) / 4 + 4;   // clist[MC], wcnt[32], flags[db_stride] bytes
    // kernel 1 (panels): shared memory
    g.ps_Aq = 0; g.ps_X0 = up4((size_t)2 * pl * pl); g.ps_M0 = up4(g.ps_X0 + (size_t)pl * n); g.ps_P00 = up4(g.ps_M0 + (size_t)pl * pl);
    g.ps_red = up4(g.ps_P00 + (size_t)pl * pl); g.ps_int = g.ps_red + 80; g.ps_doubles = g.ps_int + (ints + 1) / 2;
    // panel workspace (global, per instance): C (p x MC), V (p x MC), Xc (n x MC), clist (MC ints), meta (8)
    g.pw_C = 0; g.pw_V = (size_t)pl * g.MC; g.pw_Xc = g.pw_V + (size_t)pl * g.MC; g.pw_clist = g.pw_Xc + (size_t)n * g.MC;
    g.pw_meta = up4(g.pw_clist + (g.MC + 1) / 2); g.pw_doubles = g.pw_meta + 8;
    // kernel 2 (tiles + elimination): shared memory
    g.sm_C = 0; g.sm_V = (size_t)pl * g.MC; g.sm_Xc = g.sm_V + (size_t)pl * g.MC;
    g.sm_col = g.sm_Xc + (size_t)n * g.MC;                              // ring[3] of { colA, colAs, colW, colWs : [4][MC] }, then info[3][24]
    g.sm_red = g.sm_col + (size_t)48 * g.MC + 3 * 24;
    g.sm_int = g.sm_red + 80;
    g.smem_doubles = g.sm_int + ((size_t)g.MC + 1) / 2 + 2;
    // kept state per instance: M0 (p x p), U (p x MC), C (p x MC), L (MC x MC, column q = q-th accepted pivot), accpos (MC), meta (8)
    g.off_M0 = 0; g.off_U = up4((size_t)pl * pl); g.off_C = g.off_U + (size_t)pl * g.MC; g.off_L = g.off_C + (size_t)pl * g.MC;
    g.off_acc = g.off_L + (size_t)g.MC * g.MC; g.state_doubles = up4(g.off_acc + g.MC + 8);
    g.two_variants = g.ntiles > 352 ? 1 : 0;
    g.eligible = (p > 0 && g.MC <= 128 && g.nthreads <= 544 && g.smem_doubles * sizeof(double) <= (size_t)225 * 1024 &&
                  g.ps_doubles * sizeof(double) <= (size_t)225 * 1024) ? 1 : 0;
    return g;
}

// Kernel 1 of 2: candidate list, Pi_0^{-1} and the panels C, V, Xc of one instance, written to the global panel workspace
// (and C, U, M0 to the kept factorisation).  Small footprint (the Gauss-Jordan scratch and three p x p matrices in shared
// memory, 64 registers), so four to five instances share an SM and hide each other's pivot chains.
__global__ void __launch_bounds__(256, 4) round4_panels_kernel(Round4Params P, SchurGeom g) {
    extern __shared__ double smem[];
    const int b = blockIdx.x, n = P.n, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int p = poly_dim(n, P.cfg.polynomial_degree), pl = p;
    const int MC = g.MC;
    double* pw = P.panel_ws + (size_t)blockIdx.x * g.pw_doubles;
    double* Cs = pw + g.pw_C; double* Vs = pw + g.pw_V; double* Xc = pw + g.pw_Xc;      // panels: global (L2-resident) workspace
    double* pmeta = pw + g.pw_meta;
    double* Aq = smem + g.ps_Aq;       // Gauss-Jordan scratch [Pi_0 | I] (p x 2p)
    double* Qx = Aq + pl * pl;
    double* X0 = smem + g.ps_X0; double* M0 = smem + g.ps_M0; double* P00 = smem + g.ps_P00;
    double* red = smem + g.ps_red;
    int* clist = reinterpret_cast<int*>(smem + g.ps_int); int* wcnt = clist + MC;
    unsigned char* cflag = reinterpret_cast<unsigned char*>(wcnt + 32);

#define SCHUR_STAMPX(i) do { if (P.dbg_clock && b == 0) P.dbg_clock[(i)] = clock64(); } while (0)
#define SCHUR_STAMP(i) do { if (P.dbg_clock && b == 0 && tid == 0) P.dbg_clock[(i)] = clock64(); } while (0)
    SCHUR_STAMP(0);
    const int n_db = P.n_db[b];
    const double* sites = P.sites + (size_t)b * P.db_stride * n;
    const double* lb2 = P.lb2 + (size_t)b * n;
    const double* ub2 = P.ub2 + (size_t)b * n;
    const int* found = P.found + (size_t)b * P.found_stride;
    const int nf_ids = P.n_found[b];
    const int n_extra = P.n_extra ? P.n_extra[b] : 0;
    const double* extra = P.extra_sites ? P.extra_sites + (size_t)b * P.extra_stride * n : nullptr;
    const int N0 = nf_ids + n_extra;
    const int max_points = P.max_points;
    if (tid == 0 && P.elig) P.elig[b] = 0;
    if (tid == 0) pmeta[0] = 0.0;                  // [0] number of candidates handed to kernel 2 (0: nothing to do)
    if (!(N0 < max_points)) { if (tid == 0) { P.n_r4[b] = 0; if (P.status) P.status[b] = 0; } return; }
    if (N0 != p || n_db > MC) { if (tid == 0) P.n_r4[b] = -1; return; }      // literal kernel takes over

    // ---- candidates: results_in_box_indices(db, lb_2, ub_2, found) in ascending id order (RbfModel.jl:360).  After rounds
    // 1-3 the box-2 flags and the picked ids are already in the flag bytes of select_rounds123_kernel.
    if (P.cflags) {
        const unsigned char* cf = P.cflags + (size_t)b * P.db_stride;
        for (int id = tid; id < n_db; id += nt) { const unsigned f = cf[id]; cflag[id] = ((f & 2u) && !(f & 4u)) ? 1 : 0; }
    } else {
        for (int id = tid; id < n_db; id += nt) cflag[id] = 1;
        __syncthreads();
        for (int e = tid; e < n_db * n; e += nt) {
            const int id = e / n, k = e % n;
            const double v = sites[e];
            if (!(lb2[k] <= v && v <= ub2[k])) cflag[id] = 0;
        }
        for (int e = tid; e < nf_ids; e += nt) { const int id = found[e] - 1; if (id >= 0 && id < n_db) cflag[id] = 0; }
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
    const int TRa = (mc + 3) >> 2, MCa = TRa * 4;   // tile rows / padded candidate count actually in use
    // candidate sites, coordinate-major (Xc[k][i]); the padding columns repeat the centre (finite, never a pivot)
    for (int e = tid; e < MCa * n; e += nt) {
        const int i = e % MCa, k = e / MCa;
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
    SCHUR_STAMP(1);
    // ---- Pi_0^{-1} by Gauss-Jordan with partial pivoting on [Pi_0 | I], ONE barrier per step: the row swap and the scaling
    // of the pivot row are folded into the column updates, and warp 0 -- which always owns column kk + 1 -- finds the next
    // pivot as soon as that column is final, while the other warps are still eliminating.
    int* pivi = reinterpret_cast<int*>(red + 60);   // [2] pivot rows
    double* pivr = red + 62;                        // [2] reciprocal pivots
    if (warp == 0) {
        const int pv_ = warp_argmax_abs(Aq, 0, p, lane);
        if (lane == 0) { if (pv_ < 0) red[76] = 1.0; else { pivi[0] = pv_; pivr[0] = fast_rcp(Aq[pv_]); } }
    }
    __syncthreads();
    for (int kk = 0; kk < p; ++kk) {
        if (red[76] != 0.0) { if (tid == 0) P.n_r4[b] = -1; return; }       // Pi_0 (scaled to O(1)) is rank deficient
        const int piv = pivi[kk & 1];
        const double rp = pivr[kk & 1];
        const double* colk = Aq + kk * pl;          // multipliers: column kk as it was before the swap (nobody writes it)
        const double dkk = colk[kk];
        // warp 0 takes column kk + 1 alone, then finds the next pivot and its reciprocal; warps 1.. take quads of the other columns
        const int ncols = 2 * p - (kk + 1);
        for (int q0 = (warp == 0) ? 0 : 1 + 4 * (warp - 1); q0 < ncols; q0 += 4 * (nwarps - 1)) {
            double* col = Aq + (kk + 1 + q0) * pl;
            const int nc = (warp == 0) ? 1 : min(4, ncols - q0);
            double a_k[4], pk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) if (q < nc) { a_k[q] = col[q * pl + kk]; pk[q] = col[q * pl + piv] * rp; }   // scaled pivot-row entries
            __syncwarp();
            for (int i = lane; i < p; i += 32) {
                const double mult = (i == piv) ? dkk : colk[i];
#pragma unroll
                for (int q = 0; q < 4; ++q) if (q < nc) {
                    const double old = (i == piv) ? a_k[q] : col[q * pl + i];           // row piv receives row kk (swap)
                    col[q * pl + i] = (i == kk) ? pk[q] : fma(-mult, pk[q], old);
                }
            }
            if (warp == 0 && q0 == 0 && kk + 1 < p) {
                __syncwarp();
                const int pv_ = warp_argmax_abs(col, kk + 1, p, lane);
                if (lane == 0) { if (pv_ < 0) red[76] = 1.0; else { pivi[(kk + 1) & 1] = pv_; pivr[(kk + 1) & 1] = fast_rcp(col[pv_]); } }
            }
            if (warp == 0) break;                   // the remaining columns belong to the other warps
        }
        __syncthreads();
    }
    if (red[76] != 0.0) { if (tid == 0) P.n_r4[b] = -1; return; }
    for (int e = tid; e < p * p; e += nt) { const int r = e % p, c = e / p; M0[r + c * pl] = Qx[c + r * pl]; }   // M0 = Pi_0^{-T}
    __syncthreads();                                // Gauss-Jordan scratch is dead from here on
    double* keep = P.keep_fs ? P.keep_fs + (size_t)b * P.fs_stride : nullptr;
    if (keep) for (int e = tid; e < p * p; e += nt) keep[g.off_M0 + e] = M0[e];

    SCHUR_STAMP(2);
    // ---- panels: C = Pi_0^{-T} pi~ (Lagrange coefficients) and B = Phi(S0, candidates), one (row, 4 candidates) task per thread
    for (int t = tid; t < p * TRa; t += nt) {
        const int r = t / TRa, i4 = (t % TRa) * 4;
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
    for (int t = tid; t < p * TRa; t += nt) {
        const int r = t / TRa, i4 = (t % TRa) * 4;
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
    for (int i = tid; i < mc; i += nt) reinterpret_cast<int*>(pw + g.pw_clist)[i] = clist[i];
    if (tid == 0) { pmeta[1] = inv_s; pmeta[2] = (double)N0; pmeta[0] = (double)mc; }
    SCHUR_STAMP(3);
}

// Kernel 2 of 2: the panels come back from the workspace into shared memory, every thread computes its 4 x 4 tiles of A and W and
// the blocked elimination runs in registers.  One CTA per SM (the tiles fill the register file).
// Two launch shapes of the same code, every instance is taken by exactly one: SMALL (<= 352 tiles, i.e. <= 100 candidates: 352
// threads, so the CTA leaves ~28 k registers and 80 KB of shared memory of its SM to CTAs of other streams -- rounds 1-3, the
// panels kernel or the build of another slice of the batch) and the full shape (544 threads, <= 128 candidates).
template <bool SMALL>
__global__ void __launch_bounds__(SMALL ? 352 : 544, 1) round4_schur_kernel(Round4Params P, SchurGeom g) {
    extern __shared__ double smem[];
    const int b = blockIdx.x, n = P.n, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int p = poly_dim(n, P.cfg.polynomial_degree);
    const int MC = g.MC;
    double* Cs = smem + g.sm_C; double* Vs = smem + g.sm_V; double* Xc = smem + g.sm_Xc;
    double* ring = smem + g.sm_col; double* info = ring + 48 * MC;
    double* red = smem + g.sm_red;
    int* clist = reinterpret_cast<int*>(smem + g.sm_int);
#define SCHUR_STAMPX(i) do { if (P.dbg_clock && b == 0) P.dbg_clock[(i)] = clock64(); } while (0)
#define SCHUR_STAMP(i) do { if (P.dbg_clock && b == 0 && tid == 0) P.dbg_clock[(i)] = clock64(); } while (0)
    SCHUR_STAMP(6);
    const double* pw = P.panel_ws + (size_t)b * g.pw_doubles;
    const double* pmeta = pw + g.pw_meta;
    const int mc = (int)pmeta[0];
    if (mc == 0) return;                            // kernel 1 has already written n_r4 (0, or -1 for the literal kernel)
    const double inv_s = pmeta[1];
    const int N0 = (int)pmeta[2];
    const int max_points = P.max_points;
    int* r4 = P.r4 + (size_t)b * P.r4_stride;
    double* keep = P.keep_fs ? P.keep_fs + (size_t)b * P.fs_stride : nullptr;
    const int TRa = (mc + 3) >> 2, MCa = TRa * 4;   // tile rows / padded candidate count actually in use
    if (g.two_variants && (SMALL != (schur_tiles(TRa) <= 352))) return;      // the other launch shape's instance
    {
        const int q4 = MCa >> 2;
        const double* Cg = pw + g.pw_C; const double* Vg = pw + g.pw_V; const double* Xg = pw + g.pw_Xc;
        for (int e = tid; e < (2 * p + n) * q4; e += nt) {
            const int r = e / q4, i4 = (e % q4) * 4;
            const double* src = (r < p) ? Cg + (size_t)r * MC : ((r < 2 * p) ? Vg + (size_t)(r - p) * MC : Xg + (size_t)(r - 2 * p) * MC);
            double* dst = (r < p) ? Cs + r * MC : ((r < 2 * p) ? Vs + (r - p) * MC : Xc + (r - 2 * p) * MC);
            *reinterpret_cast<double4*>(dst + i4) = *reinterpret_cast<const double4*>(src + i4);
        }
        for (int i = tid; i < mc; i += nt) clist[i] = reinterpret_cast<const int*>(pw + g.pw_clist)[i];
    }
    // split-phase barriers of the elimination, one pair per slot of the three-deep ring of pivot blocks:
    //   barD: the diagonal tile of the block has been factorised (one arrival), barP: the whole pivot panel is published
    //   (one arrival per warp that is still alive).
    unsigned long long* barD = reinterpret_cast<unsigned long long*>(red + 64);      // [3]
    unsigned long long* barP = barD + 3;                                             // [3]
    if (tid == 0) for (int q = 0; q < 3; ++q) { mbar_init(&barD[q], 1); mbar_init(&barP[q], nwarps); }
    __syncthreads();

    SCHUR_STAMP(7);
    // ---- tiles: thread t owns tile (I, K), I >= K, of A and of W.  Tiles are numbered column by column from the LAST tile
    // column, so the tiles that are still live at pivot j (K >= j / 4) are always a prefix of the thread block.
    int tI = 0, tK = 0;
    const int ntl = schur_tiles(TRa);
    const bool has_tile = tid < ntl;
    if (has_tile) {
        int s_ = 0;
        while (((s_ + 1) * (s_ + 2)) / 2 <= tid) ++s_;
        tK = TRa - 1 - s_; tI = tK + (tid - (s_ * (s_ + 1)) / 2);
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
    SCHUR_STAMP(4);
    // ---- blocked right-looking elimination over the candidates in ascending id order, four pivots (one tile column) per block.
    //   1. the thread that owns the diagonal tile (K, K) runs the four pivot tests and eliminations inside its registers
    //      (RbfModel.jl:447-452; a rejected pivot gets a zero multiplier, so nothing downstream branches on it) -> barD
    //   2. the other tiles of tile column K finish their four pivot columns against the diagonal tile's multipliers, publish
    //      them (raw and scaled by 1/d^2) and stream the Cholesky factor rows out for mrbf_build_prepared_dev        -> barP
    //   3. every tile to the right applies the rank-4 update.  Tile column K + 1 sits in the lowest live thread ids and its
    //      diagonal tile starts step 1 of the next block as soon as its own update is done; warps without live tiles leave.
    const double thr = P.chol_thr;
    const int cap = min(max_points - N0, P.r4_stride);       // RbfModel.jl:402
    int nacc = 0, nacc_diag = -1;
    const int my_last = __shfl_sync(0xffffffffu, has_tile ? tK : -1, 0);             // lane 0 holds this warp's largest tile column
    for (int K = 0; K < TRa; ++K) {
        const int slot = K % 3;
        const unsigned par = (unsigned)((K / 3) & 1);
        const int j0 = 4 * K;
        double* cA = ring + (size_t)slot * 16 * MC; double* cAs = cA + 4 * MC; double* cW = cAs + 4 * MC; double* cWs = cW + 4 * MC;
        double* inf = info + slot * 24;              // [0..3] 1/d, [4..7] 1/d^2 (0 if rejected), [8..11] 1/(1+lev) (0 if rejected), [12] mask, [13] stop
        if (K < 40) SCHUR_STAMP(8 + 2 * K);
        if (has_tile && tK == K && tI == K) {        // ---- 1. diagonal tile
            if (K < 25) SCHUR_STAMPX(128 + 8 * K);
            int mask = 0, na = nacc;
            double rdv[4], rav[4], rwv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double dA = A[q][q], dW = W[q][q];
                const double rw = fast_rcp(dW), rs = fast_rsqrt(fmax(dA, 1e-300));      // independent: their latencies overlap
                const bool ok = (j0 + q < mc) && (na < cap) && (dA * rw > thr);      // d^2 / (1 + lev) == sigma - ||L^-1 v||^2
                const double rd = ok ? rs : 0.0;
                rdv[q] = rd; rav[q] = rd * rd; rwv[q] = ok ? rw : 0.0;
                if (ok) { mask |= 1 << q; na += 1; }
#pragma unroll
                for (int r = q + 1; r < 4; ++r)
#pragma unroll
                    for (int c = q + 1; c <= r; ++c) {
                        A[r][c] = fma(-A[r][q] * rav[q], A[c][q], A[r][c]);
                        W[r][c] = fma(-W[r][q] * rwv[q], W[c][q], W[r][c]);
                    }
            }
            if (K < 25) SCHUR_STAMPX(128 + 8 * K + 1);
            *reinterpret_cast<double4*>(inf) = make_double4(rdv[0], rdv[1], rdv[2], rdv[3]);
            *reinterpret_cast<double4*>(inf + 4) = make_double4(rav[0], rav[1], rav[2], rav[3]);
            *reinterpret_cast<double4*>(inf + 8) = make_double4(rwv[0], rwv[1], rwv[2], rwv[3]);
            *reinterpret_cast<double2*>(inf + 12) = make_double2((double)mask, (na >= cap || j0 + 4 >= mc) ? 1.0 : 0.0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {           // rows j0..j0+3 of the four pivot columns: zeros on and above the diagonal
                const double a1 = (q < 1) ? A[1][q] : 0.0, a2 = (q < 2) ? A[2][q] : 0.0, a3 = (q < 3) ? A[3][q] : 0.0;
                const double w1 = (q < 1) ? W[1][q] : 0.0, w2 = (q < 2) ? W[2][q] : 0.0, w3 = (q < 3) ? W[3][q] : 0.0;
                *reinterpret_cast<double4*>(cA + q * MC + j0) = make_double4(0.0, a1, a2, a3);
                *reinterpret_cast<double4*>(cAs + q * MC + j0) = make_double4(0.0, a1 * rav[q], a2 * rav[q], a3 * rav[q]);
                *reinterpret_cast<double4*>(cW + q * MC + j0) = make_double4(0.0, w1, w2, w3);
                *reinterpret_cast<double4*>(cWs + q * MC + j0) = make_double4(0.0, w1 * rwv[q], w2 * rwv[q], w3 * rwv[q]);
            }
            nacc_diag = na;
            mbar_arrive(&barD[slot]);
            if (K < 25) SCHUR_STAMPX(128 + 8 * K + 2);
        }
        __syncwarp();                                // panel tiles that share the diagonal tile's warp start after it, not beside it
        if (has_tile && tK == K && tI > K) {         // ---- 2. the rest of the pivot panel
            mbar_wait(&barD[slot], par);
            if (K < 25 && tI == K + 1) SCHUR_STAMPX(128 + 8 * K + 3);
            const int mask = (int)inf[12];
#pragma unroll
            for (int q = 0; q < 3; ++q)
#pragma unroll
                for (int jj = q + 1; jj < 4; ++jj) {
                    const double la = cAs[q * MC + j0 + jj], lw = cWs[q * MC + j0 + jj];
#pragma unroll
                    for (int a = 0; a < 4; ++a) { A[a][jj] = fma(-A[a][q], la, A[a][jj]); W[a][jj] = fma(-W[a][q], lw, W[a][jj]); }
                }
            int q_out = nacc;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double rd = inf[q], ra = inf[4 + q], rw = inf[8 + q];
                *reinterpret_cast<double4*>(cA + q * MC + 4 * tI) = make_double4(A[0][q], A[1][q], A[2][q], A[3][q]);
                *reinterpret_cast<double4*>(cAs + q * MC + 4 * tI) = make_double4(A[0][q] * ra, A[1][q] * ra, A[2][q] * ra, A[3][q] * ra);
                *reinterpret_cast<double4*>(cW + q * MC + 4 * tI) = make_double4(W[0][q], W[1][q], W[2][q], W[3][q]);
                *reinterpret_cast<double4*>(cWs + q * MC + 4 * tI) = make_double4(W[0][q] * rw, W[1][q] * rw, W[2][q] * rw, W[3][q] * rw);
                if (mask & (1 << q)) {
                    if (keep) *reinterpret_cast<double4*>(keep + g.off_L + (size_t)q_out * MC + 4 * tI) =
                                  make_double4(A[0][q] * rd, A[1][q] * rd, A[2][q] * rd, A[3][q] * rd);
                    q_out += 1;
                }
            }
        }
        if (K < 25 && has_tile && tK == K && tI == K + 1) SCHUR_STAMPX(128 + 8 * K + 4);
        __syncwarp();
        const bool leaving = my_last <= K;           // no tile of this warp lies to the right of tile column K
        if (lane == 0) {
            if (!leaving) mbar_arrive(&barP[slot]);
            else { mbar_arrive_drop(&barP[slot]); mbar_arrive_drop(&barP[(slot + 1) % 3]); mbar_arrive_drop(&barP[(slot + 2) % 3]); }
        }
        if (has_tile && tK == K && tI == K) {        // ids, positions and the diagonal block of the Cholesky factor (off the critical path)
            int q_out = nacc;
            const int dmask = (int)inf[12];
#pragma unroll
            for (int q = 0; q < 4; ++q) if (dmask & (1 << q)) {
                r4[q_out] = clist[j0 + q] + 1;
                if (keep) {
                    keep[g.off_acc + q_out] = (double)(j0 + q);
                    double* Lc = keep + g.off_L + (size_t)q_out * MC + j0;
                    const double rd = inf[q];
#pragma unroll
                    for (int r = 0; r < 4; ++r) if (r >= q && j0 + r < mc) Lc[r] = A[r][q] * rd;
                }
                q_out += 1;
            }
        }
        if (leaving) break;
        mbar_wait(&barP[slot], par);
        if (K < 40) SCHUR_STAMP(9 + 2 * K);
        nacc += __popc((unsigned)(int)inf[12]);
        if (inf[13] != 0.0) break;                   // capacity reached (RbfModel.jl:402) or no candidates left
        if (has_tile && tK > K) {                    // ---- 3. rank-4 update
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                {
                    const double4 i4 = *reinterpret_cast<const double4*>(cAs + q * MC + 4 * tI), k4 = *reinterpret_cast<const double4*>(cA + q * MC + 4 * tK);
                    const double ai[4] = {i4.x, i4.y, i4.z, i4.w}, ak[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int c = 0; c < 4; ++c) A[a][c] = fma(-ai[a], ak[c], A[a][c]);
                }
                {
                    const double4 i4 = *reinterpret_cast<const double4*>(cWs + q * MC + 4 * tI), k4 = *reinterpret_cast<const double4*>(cW + q * MC + 4 * tK);
                    const double wi[4] = {i4.x, i4.y, i4.z, i4.w}, wk[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int c = 0; c < 4; ++c) W[a][c] = fma(-wi[a], wk[c], W[a][c]);
                }
            }
        }
    }
    SCHUR_STAMP(5);
    // thread 0 owns the last diagonal tile: it is alive until the end and has seen every accepted pivot
    if (nacc_diag >= 0) nacc = nacc_diag;
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
    const size_t psmem = g.ps_doubles * sizeof(double), smem = g.smem_doubles * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(round4_panels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(round4_schur_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(round4_schur_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    round4_panels_kernel<<<P.B, 256, psmem, s>>>(P, g);
    if (g.two_variants) {
        round4_schur_kernel<true><<<P.B, 352, smem, s>>>(P, g);
        round4_schur_kernel<false><<<P.B, g.nthreads, smem, s>>>(P, g);
    } else if (g.nthreads <= 352) {
        round4_schur_kernel<true><<<P.B, g.nthreads, smem, s>>>(P, g);
    } else {
        round4_schur_kernel<false><<<P.B, g.nthreads, smem, s>>>(P, g);
    }
    return cudaGetLastError();
}

}  // namespace mrbf
