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
//     lev  = c_xi' (I + C_S C_S')^{-1} c_xi,   C_S = Lagrange coefficients of the accepted candidates
//     (lev = pi' (Pi' Pi)^{-1} pi is the leverage whose 1/(1+lev) is the product of the Givens cosines of
//     utilities.jl:437-448; 1 + lev is the Schur complement of I + C'C at xi -- push-through identity).
//   * so round 4 is ONE symmetric elimination on A over the candidates in ascending id order, where a rejected pivot is
//     simply skipped (RbfModel.jl:452 leaves the state untouched), with the leverages supplied beside it.  A (mc x mc,
//     mc <= 128 candidates) never touches memory: every tile thread owns one 4 x 4 tile in registers, computed directly
//     from the shared-memory panels C (Lagrange coefficients), V = B - Phi00 C / 2 and the candidate sites
//         A_ij = phi(|xi_i - xi_j|) - c_i.v_j - v_i.c_j .
//     The elimination is blocked, four pivots (one tile column) at a time: the diagonal-tile thread runs the four pivot
//     tests in registers, the tiles of that tile column publish the pivot panel, every tile to the right applies the
//     rank-4 update; the hand-overs are split-phase mbarriers.  The Cholesky factor rows (pivot column / d) are streamed
//     out for mrbf_build_prepared_dev, which finishes the model with two triangular solves per output.
//   * the leverages come from ONE extra warp per CTA that keeps M = (I + C_S C_S')^{-1} (p x p, shared memory) and hands the
//     diagonal-tile thread the 4 x 4 block G = I + C_J' M C_J of the next four candidates J; the thread eliminates G in its
//     registers beside A's diagonal tile (so every accept pattern inside the block is covered) and returns the multipliers,
//     from which the warp applies the rank-(#accepted) downdate of M.  That keeps the p x p recursion off the tile threads:
//     half the registers of an elimination on the pair (A, I + C'C), so TWO instances share an SM and hide each other's
//     pivot chains.  For p <= 32 the warp's three products per block (downdate of M, M C_J, C_J' M C_J) are mma.m8n8k4.f64
//     instructions: the warp is one serial resource per instance and its instruction count sets the pace of the elimination.
//   * two kernels: round4_prep_kernel (candidate list, Pi_0^{-1}; small footprint, several instances per SM) and
//     round4_elim_kernel (panels in shared memory, tiles, elimination; in two launch shapes: <= 100 candidates on 352 + 32
//     threads, two CTAs per SM, else 544 + 32 threads).
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

// Rows of the shared-memory panels (candidates along the row) are stored in two halves: the candidates 4t, 4t + 1 of tile t at
// [2t, 2t + 1] and 4t + 2, 4t + 3 at [H + 2t, H + 2t + 1], H = LD / 2.  A thread reads the four values of its tile with two
// 16-byte loads, and the threads of consecutive tiles touch consecutive 16-byte pieces -- no bank conflicts (32 bytes per thread
// in one piece would be a two-way conflict on every access).
__device__ __forceinline__ double4 ld4(const double* row, int t, int H) {
    const double2 a = *reinterpret_cast<const double2*>(row + 2 * t), b = *reinterpret_cast<const double2*>(row + H + 2 * t);
    return make_double4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st4(double* row, int t, int H, double4 v) {
    *reinterpret_cast<double2*>(row + 2 * t) = make_double2(v.x, v.y);
    *reinterpret_cast<double2*>(row + H + 2 * t) = make_double2(v.z, v.w);
}
__device__ __forceinline__ int off4(int i, int H) { return ((i >> 2) << 1) + (i & 1) + ((i >> 1) & 1) * H; }

// Shared-memory layout of round4_elim_kernel for leading dimension LD (the padded candidate count of the instance, <= the
// launch's maximum), in doubles; every block starts 32-byte aligned (double4 accesses).
//   phase "panels":  C (p x LD) | V (p x LD) | Xc (n x LD) || X0 (p x n) | M0 (p x p) | P00 (p x p) || fixed
//   phase "elim":    C (p x LD) | ring[3] of { colA, colAs : [4][LD] } .... || M (p x PS) | su (p x 4)  || fixed
// fixed: G blocks [3][16], info [3][28], barriers [9], clist (LD ints)
constexpr int LEV_MS = 36;      // row stride of M for the tensor-path leverage warp: rows 32 bytes apart modulo 128 (fragment loads without bank conflicts)
struct ElimLayout { int C, V, Xc, ring, X0, M0, P00, Minv, su, gs, inf, bar, clist, total, PS; };
__host__ __device__ inline ElimLayout elim_layout(int p, int n, int LD) {
    ElimLayout L;
    const int pl = p > 0 ? p : 1;
    const int pg = (pl + 7) & ~7;                    // the leverage warp walks the columns of M in groups of eight
    L.PS = pg + 1;                                   // odd row stride: a lane per row walks the columns without bank conflicts
    L.C = 0; L.V = pl * LD; L.Xc = L.V + pl * LD; L.ring = L.V;
    int regB = (pl + n) * LD; if (regB < 24 * LD) regB = 24 * LD;
    const int baseC = L.V + regB;
    L.X0 = baseC; L.M0 = (L.X0 + pl * (n | 1) + 3) & ~3; L.P00 = (L.M0 + pl * pl + 3) & ~3;     // X0: odd row stride (no bank conflicts)
    int endC1 = L.P00 + pl * pl;
    L.Minv = baseC; L.su = (L.Minv + pl * L.PS + 3) & ~3;
    int endC2 = L.su + 4 * pg;
    if (pl <= 32) { L.su = L.Minv + 32 * LEV_MS; endC2 = L.su + 3 * 128; }     // tensor-path leverage warp: M 32 x LEV_MS, u~, -u~/(1+lev), U (32 x 4 each)
    int endC = endC1 > endC2 ? endC1 : endC2; endC = (endC + 3) & ~3;
    L.gs = endC; L.inf = L.gs + 48; L.bar = L.inf + 84; L.clist = L.bar + 12;
    L.total = L.clist + (LD + 1) / 2 + 2;
    return L;
}

// the same wait for warps that have slack: sleep between the polls instead of competing for issue slots with the pivot chain
__device__ __forceinline__ void mbar_wait_relaxed(unsigned long long* bar, unsigned parity) {
    unsigned done;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(200);
    }
}

SchurGeom round4_schur_geom(int n, int p, int db_stride) {
    SchurGeom g{};
    const int pl = p > 0 ? p : 1;
    g.MC = (db_stride + 3) & ~3;
    if (g.MC < 4) g.MC = 4;
    g.TR = g.MC / 4;
    g.ntiles = schur_tiles(g.TR);
    g.two_variants = g.ntiles > 352 ? 1 : 0;
    int nt = (g.ntiles + 31) & ~31;
    if (nt < 64) nt = 64;
    g.nthreads = nt + 32;                              // tile threads + the leverage warp
    g.LD_small = g.MC < 104 ? g.MC : 104;                // 26 tile rows = 351 tiles still run in the small shape
    g.smem_small = (size_t)elim_layout(p, n, g.LD_small).total;
    g.smem_doubles = (size_t)elim_layout(p, n, g.MC).total;
    auto up4 = [](size_t v) { return (v + 3) & ~(size_t)3; };          // every block starts 32-byte aligned (double4 accesses)
    const size_t ints = (size_t)g.MC + 32 + ((size_t)db_stride + 3) / 4 + 4;   // clist[MC], wcnt[32], flags[db_stride] bytes
    // kernel 1 (prep): shared memory
    g.ps_Aq = 0; g.ps_X0 = up4((size_t)2 * pl * pl); g.ps_red = up4(g.ps_X0 + (size_t)pl * n);
    g.ps_int = g.ps_red + 80; g.ps_doubles = g.ps_int + (ints + 1) / 2;
    // hand-over workspace (global, per instance): M0 = Pi_0^{-T} (p x p), clist (MC ints), meta (8)
    g.pw_M0 = 0; g.pw_clist = up4((size_t)pl * pl);
    g.pw_meta = up4(g.pw_clist + (g.MC + 1) / 2); g.pw_doubles = g.pw_meta + 8;
    // kept state per instance: M0 (p x p), U (p x MC), C (p x MC), L (MC x MC, column q = q-th accepted pivot), accpos (MC), meta (8)
    g.off_M0 = 0; g.off_U = up4((size_t)pl * pl); g.off_C = g.off_U + (size_t)pl * g.MC; g.off_L = g.off_C + (size_t)pl * g.MC;
    g.off_acc = g.off_L + (size_t)g.MC * g.MC; g.state_doubles = up4(g.off_acc + g.MC + 8);
    g.eligible = (p > 0 && p <= 96 && g.MC <= 128 && g.ntiles <= 544 && g.smem_doubles * sizeof(double) <= (size_t)225 * 1024 &&
                  g.ps_doubles * sizeof(double) <= (size_t)225 * 1024) ? 1 : 0;
    return g;
}

// clock64 stamps of CTA 0 (MRBF_DEBUG_CLOCK=1); the elimination kernel is instantiated without them for production launches
#define SCHUR_STAMPX(i) do { if (SCHUR_DBG && P.dbg_clock && b == 0) P.dbg_clock[(i)] = clock64(); } while (0)
#define SCHUR_STAMP(i) do { if (SCHUR_DBG && P.dbg_clock && b == 0 && tid == 0) P.dbg_clock[(i)] = clock64(); } while (0)

// Kernel 1 of 2: candidate list and Pi_0^{-1} of one instance, written to the hand-over workspace.  Small footprint (the
// Gauss-Jordan scratch and the found sites in shared memory), so several instances share an SM and hide each other's pivot chains.
__global__ void __launch_bounds__(256, 4) round4_prep_kernel(Round4Params P, SchurGeom g) {
    constexpr bool SCHUR_DBG = true;
    extern __shared__ double smem[];
    const int b = blockIdx.x, n = P.n, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int p = poly_dim(n, P.cfg.polynomial_degree), pl = p;
    const int MC = g.MC;
    double* pw = P.panel_ws + (size_t)blockIdx.x * g.pw_doubles;
    double* pmeta = pw + g.pw_meta;
    double* Aq = smem + g.ps_Aq;       // Gauss-Jordan scratch [Pi_0 | I] (p x 2p)
    double* Qx = Aq + pl * pl;
    double* X0 = smem + g.ps_X0;
    double* red = smem + g.ps_red;
    int* clist = reinterpret_cast<int*>(smem + g.ps_int); int* wcnt = clist + MC;
    unsigned char* cflag = reinterpret_cast<unsigned char*>(wcnt + 32);

    SCHUR_STAMP(0);
    const bool bm = P.build_mode != 0;
    const int n_db = P.n_db[b];
    const double* sites = P.sites + (size_t)b * P.db_stride * n;
    const double* lb2 = bm ? nullptr : P.lb2 + (size_t)b * n;
    const double* ub2 = bm ? nullptr : P.ub2 + (size_t)b * n;
    const int* found = bm ? nullptr : P.found + (size_t)b * P.found_stride;
    const int nf_ids = bm ? p : P.n_found[b];
    const int n_extra = P.n_extra ? P.n_extra[b] : 0;
    const double* extra = P.extra_sites ? P.extra_sites + (size_t)b * P.extra_stride * n : nullptr;
    // hand-over from the literal kernel's prefix run (Round4Params::hyb): npre accepted round-4 points complete the found set
    const int hyb = (!bm && P.hyb) ? P.hyb[b] : 0;
    const int npre = (hyb == 1) ? P.pre_cnt[b] : 0, min_id = (hyb == 1) ? P.pre_min[b] : 0;
    const int* r4pre = P.r4 + (size_t)b * P.r4_stride;
    const int N0 = nf_ids + n_extra + npre;
    const int max_points = P.max_points;
    if (tid == 0) { pmeta[0] = 0.0; pmeta[3] = (double)npre; }      // [0] number of candidates handed to kernel 2 (0: nothing to do)
    if (tid == 0 && P.elig) P.elig[b] = 0;
    if (hyb == 2) return;                          // finished by the literal kernel: results are in place
    if (bm) {
        // build mode: found set = the first p training points, candidates = all the others; too few / too many points: general kernel
        if (n_db <= p || n_db - p > MC) { if (tid == 0) P.n_r4[b] = -1; return; }
        for (int i = tid; i < p; i += nt) P.found_out[(size_t)b * P.found_stride + i] = i + 1;
        if (tid == 0) P.n_found_out[b] = p;
    } else {
        if (!(N0 < max_points)) { if (tid == 0) { P.n_r4[b] = npre; if (P.status) P.status[b] = 0; } return; }
        if (N0 != p || n_db > MC) { if (tid == 0) P.n_r4[b] = -1; return; }      // literal kernel takes over
    }

    // ---- candidates: results_in_box_indices(db, lb_2, ub_2, found) in ascending id order (RbfModel.jl:360).  After rounds
    // 1-3 the box-2 flags and the picked ids are already in the flag bytes of select_rounds123_kernel.
    if (bm) {
        // nothing to scan
    } else if (P.cflags) {
        const unsigned char* cf = P.cflags + (size_t)b * P.db_stride;
        for (int id = tid; id < n_db; id += nt) { const unsigned f = cf[id]; cflag[id] = ((f & 2u) && !(f & 4u) && id >= min_id) ? 1 : 0; }
    } else {
        for (int id = tid; id < n_db; id += nt) cflag[id] = (id >= min_id) ? 1 : 0;
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
        X0[e] = bm ? sites[e] : ((i < nf_ids) ? sites[(size_t)(found[i] - 1) * n + k]
                                 : ((i < nf_ids + n_extra) ? extra[(size_t)(i - nf_ids) * n + k] : sites[(size_t)(r4pre[i - nf_ids - n_extra] - 1) * n + k]));
    }
    if (tid == 0) red[76] = 0.0;
    __syncthreads();
    int mc = 0;
    if (bm) {
        mc = n_db - p;
        for (int i = tid; i < mc; i += nt) clist[i] = p + i;
    } else {
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
    if (mc == 0) { if (tid == 0) { P.n_r4[b] = npre; if (P.status) P.status[b] = 0; } return; }
    if (p <= 32 && nwarps == 8) {
        // ---- Pi_0^{-1} by Gauss-Jordan on [Pi_0 | I] held in REGISTERS: lane i owns row i, warp w the columns w, w + 8, ..  The pivot
        // row of a step is never moved (its index is remembered instead of a row swap: row k of the inverse is row r_k of the right
        // half), the pivot-row entries travel by shuffle and the multiplier column through a double-buffered shared-memory column,
        // ONE barrier per step.  The warp that owns column k + 1 looks for the next pivot right after its own update.
        double a[8];
#pragma unroll
        for (int sl = 0; sl < 8; ++sl) {
            const int c = warp + 8 * sl;
            double v = 0.0;
            if (lane < p) {
                if (c < p) v = (c == 0) ? 1.0 : (X0[lane * n + c - 1] - X0[c - 1]) * inv_s;
                else if (c < 2 * p) v = (c - p == lane) ? 1.0 : 0.0;
            }
            a[sl] = v;
        }
        double* fb = Aq;                               // [2][32] multiplier columns
        double* rpb = Aq + 64;                         // [2] reciprocal pivots
        int* ib = reinterpret_cast<int*>(Aq + 66);     // [2] pivot rows (-1: singular)
        bool used = lane >= p;
        int mystep = -1;
#pragma unroll
        for (int so = 0; so < 4; ++so)               // column kk lives in slot so = kk / 8 (compile-time: the array stays in registers)
        for (int kk = 8 * so; kk < min(p, 8 * so + 8); ++kk) {
            const int par = kk & 1;
            if (warp == (kk & 7)) {
                const double v = a[so];
                const double av = used ? 0.0 : fabs(v);
                const unsigned vh = (unsigned)__double2hiint(av), vl = (unsigned)__double2loint(av);
                const unsigned mh = __reduce_max_sync(0xffffffffu, vh);
                const unsigned ml = __reduce_max_sync(0xffffffffu, vh == mh ? vl : 0u);
                const unsigned win = __ballot_sync(0xffffffffu, vh == mh && vl == ml && !used);
                const double mv = __hiloint2double((int)mh, (int)ml);
                const int r = (mv > ((bm || npre > 0) ? 1e-6 : 1e-12) && win) ? (__ffs(win) - 1) : -1;      // build mode: an ill-conditioned Pi_0 goes to the QR-based kernel
                const double pv = __shfl_sync(0xffffffffu, v, r < 0 ? 0 : r);
                fb[par * 32 + lane] = v;
                if (lane == 0) { ib[par] = r; rpb[par] = (r < 0) ? 0.0 : fast_rcp(pv); }
            }
            __syncthreads();
            const int r = ib[par];
            if (r < 0) { if (tid == 0) P.n_r4[b] = -1; return; }       // Pi_0 (scaled to O(1)) is rank deficient
            const double rp = rpb[par], f = fb[par * 32 + lane];
#pragma unroll
            for (int sl = 0; sl < 8; ++sl) {
                const double pv = __shfl_sync(0xffffffffu, a[sl], r) * rp;
                a[sl] = (lane == r) ? pv : fma(-f, pv, a[sl]);
            }
            if (lane == r) { used = true; mystep = kk; }
        }
        // row r_k of the right half is row k of Pi_0^{-1};  M0 = Pi_0^{-T}:  M0[j + k * p] = inverse[k][j]
        if (lane < p) {
#pragma unroll
            for (int sl = 0; sl < 8; ++sl) {
                const int c = warp + 8 * sl;
                if (c >= p && c < 2 * p) pw[g.pw_M0 + (c - p) + mystep * pl] = a[sl];
            }
        }
        for (int i = tid; i < mc; i += nt) reinterpret_cast<int*>(pw + g.pw_clist)[i] = clist[i];
        if (tid == 0) { pmeta[1] = inv_s; pmeta[2] = (double)N0; pmeta[0] = (double)mc; }
        SCHUR_STAMP(1);
        SCHUR_STAMP(2);
        return;
    }
    for (int e = tid; e < p * p; e += nt) {
        const int i = e % p, j = e / p;
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
    for (int e = tid; e < p * p; e += nt) { const int r = e % p, c = e / p; pw[g.pw_M0 + r + c * pl] = Qx[c + r * pl]; }   // M0 = Pi_0^{-T}
    for (int i = tid; i < mc; i += nt) reinterpret_cast<int*>(pw + g.pw_clist)[i] = clist[i];
    if (tid == 0) { pmeta[1] = inv_s; pmeta[2] = (double)N0; pmeta[0] = (double)mc; }
    SCHUR_STAMP(2);
}

// Sum of 16 per-lane values over the warp with 32 shuffles: each exchange step keeps one half of the values and sends the other.
// Afterwards lane l holds the total of value  8 * bit4(l) + 4 * bit3(l) + 2 * bit2(l) + bit1(l).
__device__ __forceinline__ double warp_sum16(double (&v)[16], int lane) {
#pragma unroll
    for (int half = 8, m = 16; half >= 1; half >>= 1, m >>= 1) {
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const double keep = up ? v[i + half] : v[i];
            const double send = up ? v[i] : v[i + half];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
        }
    }
    return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// Results of pivot block K, written from the published panel (ring slot) by threads that have slack -- never by the warp that
// carries the pivot chain: ids of the accepted candidates, their positions, and the rows of the Cholesky factor (pivot column / d)
// for mrbf_build_prepared_dev.  Worker w of nw takes the (column q, tile row I) items w, w + nw, ..
__device__ __forceinline__ void store_pivot_block(const double* cA, const double* inf, int K, int TRa, int LD, int MC, int nacc0, int* r4,
                                                  const int* clist, double* keep, size_t off_L, size_t off_acc, int w, int nw) {
    const int mask = (int)inf[12];
    if (mask == 0) return;
    const int j0 = 4 * K, items = keep ? 4 * (TRa - K) : 4;
    for (int e = w; e < items; e += nw) {
        const int q = e & 3, I = K + (e >> 2);
        if (!(mask & (1 << q))) continue;
        const int q_out = nacc0 + __popc((unsigned)(mask & ((1 << q) - 1)));
        if (I == K) r4[q_out] = clist[j0 + q] + 1;
        if (keep) {
            if (I == K) keep[off_acc + q_out] = (double)(j0 + q);
            const double rd = sqrt(inf[4 + q]);      // 1/d from 1/d^2
            const double4 v = ld4(cA + q * LD, I, LD >> 1);
            *reinterpret_cast<double4*>(keep + off_L + (size_t)q_out * MC + 4 * I) = make_double4(v.x * rd, v.y * rd, v.z * rd, v.w * rd);
        }
    }
}

// One pass of the leverage warp over the rows of M it owns (lane l: rows l, l + 32, ..): downdate  M -= sum_q av_q u_q'  (upd) and
// U = M C_J, eight columns at a time so that the loads of a group are in flight together.  Columns p .. PS-2 of M are zero padding.
template <int RPL>
__device__ __forceinline__ void lev_pass(double* __restrict__ Minv, const double* __restrict__ su, const double* __restrict__ Cs, int K, int LD, int PS,
                                         int p, int lane, bool upd, const double (&av)[RPL][4], double (&U4)[RPL][4]) {
    const int pg = (p + 7) & ~7;
#pragma unroll
    for (int rr = 0; rr < RPL; ++rr) {
        const int r = lane + 32 * rr;
        double u0 = 0.0, u1 = 0.0, u2 = 0.0, u3 = 0.0;
        if (r < p) {
            double* mrow = Minv + r * PS;
            const double a0 = av[rr][0], a1 = av[rr][1], a2 = av[rr][2], a3 = av[rr][3];
            for (int c0 = 0; c0 < pg; c0 += 8) {
                double m[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) m[i] = mrow[c0 + i];
                if (upd) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        double4 s[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) s[i] = *reinterpret_cast<const double4*>(su + 4 * (c0 + 4 * h + i));
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const double a = fma(-a1, s[i].y, fma(-a0, s[i].x, m[4 * h + i]));
                            const double t = fma(a3, s[i].w, a2 * s[i].z);
                            m[4 * h + i] = a - t;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) mrow[c0 + i] = m[i];
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    double4 cj[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) { const int c = min(c0 + 4 * h + i, p - 1); cj[i] = ld4(Cs + c * LD, K, LD >> 1); }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        u0 = fma(m[4 * h + i], cj[i].x, u0); u1 = fma(m[4 * h + i], cj[i].y, u1);
                        u2 = fma(m[4 * h + i], cj[i].z, u2); u3 = fma(m[4 * h + i], cj[i].w, u3);
                    }
                }
            }
        }
        U4[rr][0] = u0; U4[rr][1] = u1; U4[rr][2] = u2; U4[rr][3] = u3;
    }
}

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {      // D (8 x 8) += A (8 x 4) B (4 x 8), FP64 tensor path
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

// Kernel 2 of 2: the panels C, V and the candidate sites are computed into shared memory, every tile thread computes its
// 4 x 4 tile of A and the blocked elimination runs in registers; the last warp supplies the leverage blocks (header comment).
// Two launch shapes of the same code, every instance is taken by exactly one: SMALL (<= 352 tiles, i.e. <= 100 candidates:
// 352 + 32 threads, <= 80 registers and < 113 KB of shared memory, so two instances share an SM) and the full shape
// (544 + 32 threads, <= 128 candidates).  RPL = rows of M per lane of the leverage warp (p <= 32 RPL).
template <bool SMALL, int RPL, bool SCHUR_DBG>
__global__ void __launch_bounds__(SMALL ? 384 : 576, SMALL ? 2 : 1) round4_elim_kernel(Round4Params P, SchurGeom g) {
    extern __shared__ double smem[];
    const int b = blockIdx.x, n = P.n, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int p = poly_dim(n, P.cfg.polynomial_degree);
    const int MC = g.MC;
    SCHUR_STAMP(6);
    const double* pw = P.panel_ws + (size_t)b * g.pw_doubles;
    const double* pmeta = pw + g.pw_meta;
    const int mc = (int)pmeta[0];
    if (mc == 0) return;                            // kernel 1 has already written n_r4 (0, or -1 for the literal kernel)
    const double inv_s = pmeta[1];
    const int N0 = (int)pmeta[2];
    const int npre = (int)pmeta[3];                 // round-4 points the literal kernel accepted before the hand-over (part of S0 here)
    const int max_points = P.max_points;
    int* r4 = P.r4 + (size_t)b * P.r4_stride + npre;
    double* keep = (P.keep_fs && npre == 0) ? P.keep_fs + (size_t)b * P.fs_stride : nullptr;    // a handed-over instance keeps no factorisation (elig stays 0)
    const int TRa = (mc + 3) >> 2, MCa = TRa * 4;   // tile rows / padded candidate count actually in use
    if (g.two_variants && (SMALL != (schur_tiles(TRa) <= 352))) return;      // the other launch shape's instance
    const int LD = MCa, H = LD >> 1;
    const bool bm = P.build_mode != 0;
    RadFn rf = P.rf;
    if (P.shape_arr) {                              // per-instance shape parameter (NaN: the configuration's default)
        double a_ = P.alpha_default; const double sp = P.shape_arr[b]; if (sp == sp) a_ = sp;
        rf.alpha2 = a_ * a_;
    }
    const ElimLayout L = elim_layout(p, n, LD);
    double* Cs = smem + L.C; double* Vs = smem + L.V; double* Xc = smem + L.Xc;
    double* X0 = smem + L.X0; double* M0 = smem + L.M0; double* P00 = smem + L.P00;
    double* ring = smem + L.ring; double* Minv = smem + L.Minv; double* su = smem + L.su;
    double* gs = smem + L.gs; double* info = smem + L.inf;
    unsigned long long* barD = reinterpret_cast<unsigned long long*>(smem + L.bar);      // [3] diagonal tile factorised
    unsigned long long* barP = barD + 3;                                                 // [3] pivot panel published
    unsigned long long* barG = barP + 3;                                                 // [3] leverage block ready
    int* clist = reinterpret_cast<int*>(smem + L.clist);
    const int PS = L.PS;
    const int nx = n | 1;                           // row stride of X0

    // ---- found sites, Pi_0^{-T}, candidate list and candidate sites (coordinate-major; the padding columns repeat the
    // centre: finite, never a pivot)
    {
        const double* sites = P.sites + (size_t)b * P.db_stride * n;
        const int* found = bm ? nullptr : P.found + (size_t)b * P.found_stride;
        const int nf_ids = bm ? p : P.n_found[b];
        const double* extra = P.extra_sites ? P.extra_sites + (size_t)b * P.extra_stride * n : nullptr;
        for (int e = tid; e < p * n; e += nt) {
            const int i = e / n, k = e % n;
            X0[i * nx + k] = bm ? sites[e] : ((i < nf_ids) ? sites[(size_t)(found[i] - 1) * n + k]
                                              : ((i < p - npre) ? extra[(size_t)(i - nf_ids) * n + k] : sites[(size_t)(r4[i - p] - 1) * n + k]));   // r4[-npre .. -1]: the literal kernel's prefix
        }
        for (int e = tid; e < p * p; e += nt) { const double v = pw[g.pw_M0 + e]; M0[e] = v; if (keep) keep[g.off_M0 + e] = v; }
        const int* cl = reinterpret_cast<const int*>(pw + g.pw_clist);
        for (int i = tid; i < mc; i += nt) clist[i] = cl[i];
        for (int e = tid; e < mc * n; e += nt) {
            const int i = e / n, k = e % n;
            Xc[k * LD + off4(i, H)] = sites[(size_t)cl[i] * n + k];
        }
        if (tid == 0) for (int q = 0; q < 3; ++q) { mbar_init(&barD[q], 1); mbar_init(&barP[q], nwarps - 1); mbar_init(&barG[q], 1); }
        __syncthreads();
        for (int e = tid; e < (MCa - mc) * n; e += nt) { const int i = mc + e / n, k = e % n; Xc[k * LD + off4(i, H)] = X0[k]; }
        for (int e = tid; e < p * p; e += nt) {
            const int i = e % p, j = e / p;
            double r2 = 0.0;
            for (int k = 0; k < n; ++k) { const double d = X0[i * nx + k] - X0[j * nx + k]; r2 = fma(d, d, r2); }
            P00[i + j * p] = rad_phi(rf, r2);
        }
        __syncthreads();
    }
    SCHUR_STAMP(100);
    // ---- panels: C = Pi_0^{-T} pi~ (Lagrange coefficients) and B = Phi(S0, candidates), one (row, 4 candidates) task per thread
    for (int t = tid; t < p * TRa; t += nt) {
        const int r = t / TRa, i4 = (t % TRa) * 4;
        double c0 = M0[r], c1 = c0, c2 = c0, c3 = c0;            // pi~[0] = 1
        double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
        for (int k = 0; k < n; ++k) {
            const double4 x = ld4(Xc + k * LD, i4 >> 2, H);
            if (k + 1 < p) {
                const double mv = M0[r + (k + 1) * p] * inv_s, xc = X0[k];
                c0 = fma(mv, x.x - xc, c0); c1 = fma(mv, x.y - xc, c1); c2 = fma(mv, x.z - xc, c2); c3 = fma(mv, x.w - xc, c3);
            }
            const double xr = X0[r * nx + k];
            double e_;
            e_ = x.x - xr; d0 = fma(e_, e_, d0); e_ = x.y - xr; d1 = fma(e_, e_, d1);
            e_ = x.z - xr; d2 = fma(e_, e_, d2); e_ = x.w - xr; d3 = fma(e_, e_, d3);
        }
        st4(Cs + r * LD, i4 >> 2, H, make_double4(c0, c1, c2, c3));
        st4(Vs + r * LD, i4 >> 2, H, make_double4(rad_phi(rf, d0), rad_phi(rf, d1), rad_phi(rf, d2), rad_phi(rf, d3)));
        if (keep) *reinterpret_cast<double4*>(keep + g.off_C + (size_t)r * MC + i4) = make_double4(c0, c1, c2, c3);
    }
    __syncthreads();
    // ---- U = B - Phi00 C (kept for the build), V = B - Phi00 C / 2 (in place of B)
    for (int t = tid; t < p * TRa; t += nt) {
        const int r = t / TRa, i4 = (t % TRa) * 4;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        for (int c = 0; c < p; ++c) {
            const double pv = P00[r + c * p];
            const double4 cc = ld4(Cs + c * LD, i4 >> 2, H);
            s0 = fma(pv, cc.x, s0); s1 = fma(pv, cc.y, s1); s2 = fma(pv, cc.z, s2); s3 = fma(pv, cc.w, s3);
        }
        const double4 bb = ld4(Vs + r * LD, i4 >> 2, H);
        if (keep) *reinterpret_cast<double4*>(keep + g.off_U + (size_t)r * MC + i4) = make_double4(bb.x - s0, bb.y - s1, bb.z - s2, bb.w - s3);
        st4(Vs + r * LD, i4 >> 2, H, make_double4(fma(-0.5, s0, bb.x), fma(-0.5, s1, bb.y), fma(-0.5, s2, bb.z), fma(-0.5, s3, bb.w)));
    }
    __syncthreads();                                // X0, M0, P00 are dead from here on

    SCHUR_STAMP(7);
    // ---- tiles: thread t owns tile (I, K), I >= K, of A.  Tiles are numbered column by column from the LAST tile
    // column, so the tiles that are still live at pivot j (K >= j / 4) are always a prefix of the thread block.
    const bool lev_warp = warp == nwarps - 1;
    int tI = 0, tK = 0;
    const int ntl = schur_tiles(TRa);
    const bool has_tile = tid < ntl;                // ntl <= nt - 32: never a thread of the leverage warp
    if (has_tile) {
        int s_ = 0;
        while (((s_ + 1) * (s_ + 2)) / 2 <= tid) ++s_;
        tK = TRa - 1 - s_; tI = tK + (tid - (s_ * (s_ + 1)) / 2);
    }
    double A[4][4];
    if (has_tile) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) A[a][c] = 0.0;
        for (int k = 0; k < n; ++k) {
            const double4 vi = ld4(Xc + k * LD, tI, H);
            const double4 vk = ld4(Xc + k * LD, tK, H);
            const double ri[4] = {vi.x, vi.y, vi.z, vi.w}, rk[4] = {vk.x, vk.y, vk.z, vk.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) { const double d = ri[a] - rk[c]; A[a][c] = fma(d, d, A[a][c]); }
        }
        SCHUR_STAMP(101);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) A[a][c] = rad_phi(rf, A[a][c]);
        SCHUR_STAMP(102);
        for (int r = 0; r < p; ++r) {
            const double4 a4 = ld4(Cs + r * LD, tI, H), b4 = ld4(Cs + r * LD, tK, H);
            const double4 c4 = ld4(Vs + r * LD, tI, H), d4 = ld4(Vs + r * LD, tK, H);
            const double ci[4] = {a4.x, a4.y, a4.z, a4.w}, ck[4] = {b4.x, b4.y, b4.z, b4.w};
            const double vi[4] = {c4.x, c4.y, c4.z, c4.w}, vk[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    A[a][c] = fma(-ci[a], vk[c], A[a][c]);
                    A[a][c] = fma(-vi[a], ck[c], A[a][c]);
                }
        }
    } else if (lev_warp) {
        if (RPL == 1) {                             // tensor-path layout: M 32 x LEV_MS (zero beyond p), u~ / -u~/(1+lev) / U 32 x 4
            for (int e = lane; e < 32 * LEV_MS; e += 32) Minv[e] = (e / LEV_MS == e % LEV_MS && e / LEV_MS < p) ? 1.0 : 0.0;
            for (int e = lane; e < 3 * 128; e += 32) su[e] = 0.0;
        } else {
            for (int e = lane; e < p * PS; e += 32) Minv[e] = (e / PS == e % PS) ? 1.0 : 0.0;      // no candidate accepted yet: M = I
            for (int e = lane; e < 4 * (PS - 1); e += 32) su[e] = 0.0;
        }
    }
    __syncthreads();                                // V, Xc are dead from here on: the ring of pivot panels takes their place
    SCHUR_STAMP(4);

    const double thr = P.chol_thr;
    const int cap = min(max_points - N0, P.r4_stride - npre);       // RbfModel.jl:402

    if (lev_warp) {
        // ---- leverage warp.  Block K: (a) downdate M by the candidates accepted in block K - 1 (multipliers from the diagonal-tile
        // thread), (b) U = M C_J, (c) G = I + C_J' U  -> diagonal-tile thread.  Lane l owns rows l, l + 32, .. of M.
        double U4[RPL][4];
        int nacc_l = 0;                              // accepted before block K - 1
        for (int K = 0; K <= TRa; ++K) {
            const int slot = K % 3;
            bool upd = false;
            double av[RPL][4];
#pragma unroll
            for (int rr = 0; rr < RPL; ++rr) { av[rr][0] = 0.0; av[rr][1] = 0.0; av[rr][2] = 0.0; av[rr][3] = 0.0; }
            if (K > 0) {
                const int ps = (K - 1) % 3;
                mbar_wait(&barD[ps], (unsigned)(((K - 1) / 3) & 1));
                if (K < 25 && lane == 0) SCHUR_STAMPX(340 + 4 * K);
                const double* pin = info + ps * 28;
                const bool mine = (schur_tiles(max(TRa - (K - 1) - 2, 0)) >> 5) == 0;      // nobody else writes the results of block K - 1
                if (pin[13] != 0.0) {               // capacity reached or last block: no further leverage is needed
                    if (mine) {
                        mbar_wait(&barP[ps], (unsigned)(((K - 1) / 3) & 1));
                        store_pivot_block(ring + (size_t)ps * 8 * LD, pin, K - 1, TRa, LD, MC, nacc_l, r4, clist, keep, g.off_L, g.off_acc, lane, 32);
                    }
                    break;
                }
                if (!bm && (int)pin[12] != 0) {
                    upd = true;
                    const double4 rw = *reinterpret_cast<const double4*>(pin + 8);
                    const double l10 = pin[14], l20 = pin[15], l21 = pin[16], l30 = pin[17], l31 = pin[18], l32 = pin[19];
                    if (RPL == 1) {                 // tensor path: U of the previous block comes from shared memory, lane = row
                        const double4 Uv = *reinterpret_cast<const double4*>(su + 256 + 4 * lane);
                        const double u0 = Uv.x;
                        const double u1 = fma(-u0, l10, Uv.y);
                        const double u2 = fma(-u1, l21, fma(-u0, l20, Uv.z));
                        const double u3 = fma(-u2, l32, fma(-u1, l31, fma(-u0, l30, Uv.w)));
                        *reinterpret_cast<double4*>(su + 4 * lane) = make_double4(u0, u1, u2, u3);
                        *reinterpret_cast<double4*>(su + 128 + 4 * lane) = make_double4(-rw.x * u0, -rw.y * u1, -rw.z * u2, -rw.w * u3);
                    } else
#pragma unroll
                    for (int rr = 0; rr < RPL; ++rr) {
                        const int r = lane + 32 * rr;
                        const double u0 = U4[rr][0];
                        const double u1 = fma(-u0, l10, U4[rr][1]);
                        const double u2 = fma(-u1, l21, fma(-u0, l20, U4[rr][2]));
                        const double u3 = fma(-u2, l32, fma(-u1, l31, fma(-u0, l30, U4[rr][3])));
                        if (r < p) *reinterpret_cast<double4*>(su + 4 * r) = make_double4(u0, u1, u2, u3);
                        av[rr][0] = rw.x * u0; av[rr][1] = rw.y * u1; av[rr][2] = rw.z * u2; av[rr][3] = rw.w * u3;
                    }
                    __syncwarp();
                }
            }
            if (bm) {                               // no leverages in build mode: only the result stores of the late blocks
                if (K > 0) {
                    const int ps = (K - 1) % 3;
                    const double* pin = info + ps * 28;
                    if ((schur_tiles(max(TRa - (K - 1) - 2, 0)) >> 5) == 0) {
                        mbar_wait(&barP[ps], (unsigned)(((K - 1) / 3) & 1));
                        store_pivot_block(ring + (size_t)ps * 8 * LD, pin, K - 1, TRa, LD, MC, nacc_l, r4, clist, keep, g.off_L, g.off_acc, lane, 32);
                    }
                    nacc_l += __popc((unsigned)(int)pin[12]);
                }
                continue;
            }
            if (K < 25 && lane == 0) SCHUR_STAMPX(341 + 4 * K);
            if (RPL == 1) {
                // ---- p <= 32: the three products of the block on the FP64 tensor path (mma.m8n8k4): a tenth of the instructions of the
                // lane-per-row pass below, which made this warp -- one serial resource per instance -- the pace setter of the elimination
                //   M -= (u~ / (1 + lev)) u~'     16 tiles, accumulator fragments loaded from / stored to shared memory
                //   U  = M C_J                    32 x 8 (4 used), A fragments of M from shared memory
                //   G  = I + C_J' U               8 x 8 (4 x 4 used), U through shared memory as the B operand
                const int fg = lane >> 2, ft = lane & 3;
                double* ub = su; double* nb = su + 128; double* Ub = su + 256;
                if (upd) {                          // (u~ and -u~ / (1 + lev) were written below, before this point, by lane = row)
                    double af[4], bf[4];
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) { af[mt] = nb[(8 * mt + fg) * 4 + ft]; bf[mt] = ub[(8 * mt + fg) * 4 + ft]; }
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                        for (int nt_ = 0; nt_ < 4; ++nt_) {
                            double2* mp = reinterpret_cast<double2*>(Minv + (8 * mt + fg) * LEV_MS + 8 * nt_ + 2 * ft);
                            const double2 c2 = *mp;
                            double d[2] = {c2.x, c2.y};
                            dmma884(d, af[mt], bf[nt_]);
                            *mp = make_double2(d[0], d[1]);
                        }
                    __syncwarp();
                }
                double cf[8];                       // C[4 ks + ft][j0 + fg]: B fragment of U = M C_J and A fragment of G = C_J' U
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) { const int k = 4 * ks + ft; cf[ks] = (fg < 4 && k < p) ? Cs[k * LD + off4(4 * K + fg, H)] : 0.0; }
                double Ua[4][2];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) { Ua[mt][0] = 0.0; Ua[mt][1] = 0.0; }
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) dmma884(Ua[mt], Minv[(8 * mt + fg) * LEV_MS + 4 * ks + ft], cf[ks]);
                if (ft < 2) {
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) *reinterpret_cast<double2*>(Ub + (8 * mt + fg) * 4 + 2 * ft) = make_double2(Ua[mt][0], Ua[mt][1]);
                }
                __syncwarp();
                double Gd[2] = {0.0, 0.0};
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) dmma884(Gd, cf[ks], (fg < 4) ? Ub[(4 * ks + ft) * 4 + fg] : 0.0);
                if (fg < 4 && ft < 2) {             // G(a = fg, b = 2 ft + e), lower triangle, index a (a + 1) / 2 + b
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int gb = 2 * ft + e;
                        if (gb <= fg) gs[slot * 16 + (fg * (fg + 1)) / 2 + gb] = Gd[e] + (gb == fg ? 1.0 : 0.0);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&barG[slot]);
                if (K < 25 && lane == 0) SCHUR_STAMPX(343 + 4 * K);
                if (K > 0) {                        // results of block K - 1 (the leverage block of K is on its way: this is idle time)
                    const int ps = (K - 1) % 3;
                    const double* pin = info + ps * 28;
                    if ((schur_tiles(max(TRa - (K - 1) - 2, 0)) >> 5) == 0) {
                        mbar_wait(&barP[ps], (unsigned)(((K - 1) / 3) & 1));
                        store_pivot_block(ring + (size_t)ps * 8 * LD, pin, K - 1, TRa, LD, MC, nacc_l, r4, clist, keep, g.off_L, g.off_acc, lane, 32);
                    }
                    nacc_l += __popc((unsigned)(int)pin[12]);
                }
                continue;
            }
            lev_pass<RPL>(Minv, su, Cs, K, LD, PS, p, lane, upd, av, U4);
            if (K < 25 && lane == 0) SCHUR_STAMPX(342 + 4 * K);
            double v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.0;
#pragma unroll
            for (int rr = 0; rr < RPL; ++rr) {
                const int r = lane + 32 * rr;
                if (r < p) {
                    const double4 cr = ld4(Cs + r * LD, K, H);
                    v[0] = fma(cr.x, U4[rr][0], v[0]);
                    v[1] = fma(cr.y, U4[rr][0], v[1]); v[2] = fma(cr.y, U4[rr][1], v[2]);
                    v[3] = fma(cr.z, U4[rr][0], v[3]); v[4] = fma(cr.z, U4[rr][1], v[4]); v[5] = fma(cr.z, U4[rr][2], v[5]);
                    v[6] = fma(cr.w, U4[rr][0], v[6]); v[7] = fma(cr.w, U4[rr][1], v[7]); v[8] = fma(cr.w, U4[rr][2], v[8]);
                    v[9] = fma(cr.w, U4[rr][3], v[9]);
                }
            }
            const double tot = warp_sum16(v, lane);
            const int idx = lane >> 1;              // value held by this lane (pairs of lanes hold the same one)
            if ((lane & 1) == 0 && idx < 10) gs[slot * 16 + idx] = tot + ((idx == 0 || idx == 2 || idx == 5 || idx == 9) ? 1.0 : 0.0);
            __syncwarp();
            if (lane == 0) mbar_arrive(&barG[slot]);
            if (K < 25 && lane == 0) SCHUR_STAMPX(343 + 4 * K);
            if (K > 0) {                            // results of block K - 1 (the leverage block of K is on its way: this is idle time)
                const int ps = (K - 1) % 3;
                const double* pin = info + ps * 28;
                if ((schur_tiles(max(TRa - (K - 1) - 2, 0)) >> 5) == 0) {
                    mbar_wait(&barP[ps], (unsigned)(((K - 1) / 3) & 1));
                    store_pivot_block(ring + (size_t)ps * 8 * LD, pin, K - 1, TRa, LD, MC, nacc_l, r4, clist, keep, g.off_L, g.off_acc, lane, 32);
                }
                nacc_l += __popc((unsigned)(int)pin[12]);
            }
        }
        return;
    }

    // ---- blocked right-looking elimination over the candidates in ascending id order, four pivots (one tile column) per block.
    //   1. the thread that owns the diagonal tile (K, K) runs the four pivot tests and eliminations inside its registers, on its
    //      tile of A and on the leverage block G (RbfModel.jl:447-452; a rejected pivot gets a zero multiplier, so nothing
    //      downstream branches on it) -> barD
    //   2. the other tiles of tile column K finish their four pivot columns against the diagonal tile's multipliers, publish
    //      them (raw and scaled by 1/d^2) and stream the Cholesky factor rows out for mrbf_build_prepared_dev        -> barP
    //   3. every tile to the right applies the rank-4 update.  Tile column K + 1 sits in the lowest live thread ids and its
    //      diagonal tile starts step 1 of the next block as soon as its own update is done; warps without live tiles leave.
    int nacc = 0, nacc_diag = -1;
    bool bm_failed = false;
    const int my_last = __shfl_sync(0xffffffffu, has_tile ? tK : -1, 0);             // lane 0 holds this warp's largest tile column
    for (int K = 0; K < TRa; ++K) {
        const int slot = K % 3;
        const unsigned par = (unsigned)((K / 3) & 1);
        const int j0 = 4 * K;
        double* cA = ring + (size_t)slot * 8 * LD; double* cAs = cA + 4 * LD;
        double* inf = info + slot * 28;              // [4..7] 1/d^2 (0 if rejected), [8..11] 1/(1+lev) (0 if rejected), [12] mask, [13] stop,
                                                     // [14..19] multipliers of the leverage block, [20..25] multipliers of A's diagonal tile
        if (K < 40) SCHUR_STAMP(8 + 2 * K);
        if (has_tile && tK == K && tI == K) {        // ---- 1. diagonal tile
            if (K < 25) SCHUR_STAMPX(128 + 8 * K);
            double W[4][4];
            if (bm) {                               // plain Cholesky: no leverage, every candidate must be accepted
                W[0][0] = 1.0; W[1][0] = 0.0; W[1][1] = 1.0; W[2][0] = 0.0; W[2][1] = 0.0; W[2][2] = 1.0;
                W[3][0] = 0.0; W[3][1] = 0.0; W[3][2] = 0.0; W[3][3] = 1.0;
            } else {
                mbar_wait(&barG[slot], par);
                const double* gq = gs + slot * 16;
                W[0][0] = gq[0]; W[1][0] = gq[1]; W[1][1] = gq[2]; W[2][0] = gq[3]; W[2][1] = gq[4]; W[2][2] = gq[5];
                W[3][0] = gq[6]; W[3][1] = gq[7]; W[3][2] = gq[8]; W[3][3] = gq[9];
            }
            int mask = 0, na = nacc;
            bool failed = false;
            double rav[4], rwv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double dA = A[q][q], dW = W[q][q];
                const double rw = fast_rcp(dW), ra = fast_rcp(fmax(dA, 1e-300));        // independent: their latencies overlap
                const bool ok = (j0 + q < mc) && (na < cap) && (dA * rw > thr);      // d^2 / (1 + lev) == sigma - ||L^-1 v||^2
                rav[q] = ok ? ra : 0.0; rwv[q] = ok ? rw : 0.0;
                if (ok) { mask |= 1 << q; na += 1; }
                else if (bm && j0 + q < mc) failed = true;       // reduced kernel matrix not positive definite: the general kernel reports it
#pragma unroll
                for (int r = q + 1; r < 4; ++r)
#pragma unroll
                    for (int c = q + 1; c <= r; ++c) {
                        A[r][c] = fma(-A[r][q] * rav[q], A[c][q], A[r][c]);
                        W[r][c] = fma(-W[r][q] * rwv[q], W[c][q], W[r][c]);
                    }
            }
            if (K < 25) SCHUR_STAMPX(128 + 8 * K + 1);
            // what the panel tiles and the leverage warp need: 1/d^2, the accept mask, the multipliers of the diagonal blocks
            *reinterpret_cast<double4*>(inf + 4) = make_double4(rav[0], rav[1], rav[2], rav[3]);
            *reinterpret_cast<double4*>(inf + 8) = make_double4(rwv[0], rwv[1], rwv[2], rwv[3]);
            *reinterpret_cast<double2*>(inf + 12) = make_double2((double)mask, (na >= cap || j0 + 4 >= mc || failed) ? 1.0 : 0.0);
            inf[26] = failed ? 1.0 : 0.0;
            *reinterpret_cast<double2*>(inf + 14) = make_double2(W[1][0] * rwv[0], W[2][0] * rwv[0]);
            *reinterpret_cast<double4*>(inf + 16) = make_double4(W[2][1] * rwv[1], W[3][0] * rwv[0], W[3][1] * rwv[1], W[3][2] * rwv[2]);
            *reinterpret_cast<double4*>(inf + 20) = make_double4(A[1][0] * rav[0], A[2][0] * rav[0], A[2][1] * rav[1], A[3][0] * rav[0]);
            *reinterpret_cast<double2*>(inf + 24) = make_double2(A[3][1] * rav[1], A[3][2] * rav[2]);
#pragma unroll
            for (int q = 0; q < 4; ++q)             // rows j0..j0+3 of the four pivot columns (zeros above the diagonal) for store_pivot_block
                st4(cA + q * LD, K, H, make_double4(q <= 0 ? A[0][q] : 0.0, q <= 1 ? A[1][q] : 0.0, q <= 2 ? A[2][q] : 0.0, A[3][q]));
            nacc_diag = na;
            mbar_arrive(&barD[slot]);
            if (K < 25) SCHUR_STAMPX(128 + 8 * K + 2);
        }
        __syncwarp();                                // panel tiles that share the diagonal tile's warp start after it, not beside it
        if (has_tile && tK == K && tI > K) {         // ---- 2. the rest of the pivot panel
            mbar_wait(&barD[slot], par);
            if (K < 25 && tI == K + 1) SCHUR_STAMPX(128 + 8 * K + 3);
            {
                const double4 m4 = *reinterpret_cast<const double4*>(inf + 20); const double2 m2 = *reinterpret_cast<const double2*>(inf + 24);
                const double l10 = m4.x, l20 = m4.y, l21 = m4.z, l30 = m4.w, l31 = m2.x, l32 = m2.y;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    A[a][1] = fma(-A[a][0], l10, A[a][1]);
                    A[a][2] = fma(-A[a][1], l21, fma(-A[a][0], l20, A[a][2]));
                    A[a][3] = fma(-A[a][2], l32, fma(-A[a][1], l31, fma(-A[a][0], l30, A[a][3])));
                }
            }
            const double4 ra4 = *reinterpret_cast<const double4*>(inf + 4);
            const double ra[4] = {ra4.x, ra4.y, ra4.z, ra4.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                st4(cA + q * LD, tI, H, make_double4(A[0][q], A[1][q], A[2][q], A[3][q]));
                st4(cAs + q * LD, tI, H, make_double4(A[0][q] * ra[q], A[1][q] * ra[q], A[2][q] * ra[q], A[3][q] * ra[q]));
            }
        }
        if (K < 25 && has_tile && tK == K && tI == K + 1) SCHUR_STAMPX(128 + 8 * K + 4);
        __syncwarp();
        const bool leaving = my_last <= K;           // no tile of this warp lies to the right of tile column K
        if (lane == 0) {
            if (K < 25 && warp < 12) SCHUR_STAMPX(600 + 12 * K + warp);
            if (!leaving) mbar_arrive(&barP[slot]);
            else { mbar_arrive_drop(&barP[slot]); mbar_arrive_drop(&barP[(slot + 1) % 3]); mbar_arrive_drop(&barP[(slot + 2) % 3]); }
        }
        if (leaving) break;
        if (__any_sync(0xffffffffu, has_tile && tK == K + 1)) mbar_wait(&barP[slot], par);      // next tile column: on the pivot chain
        else mbar_wait_relaxed(&barP[slot], par);
        if (K < 40) SCHUR_STAMP(9 + 2 * K);
        if (K < 25 && has_tile && tK == K + 1 && tI == K + 1) SCHUR_STAMPX(128 + 8 * K + 5);
        {   // results of this block: by the warps whose tiles all lie right of tile column K + 1 (they have arrived long ago and have
            // slack), or by the leverage warp when no such warp is left
            const int nfw = schur_tiles(max(TRa - K - 2, 0)) >> 5;
            if (warp < nfw) store_pivot_block(cA, inf, K, TRa, LD, MC, nacc, r4, clist, keep, g.off_L, g.off_acc, tid, nfw * 32);
        }
        nacc += __popc((unsigned)(int)inf[12]);
        if (inf[26] != 0.0) bm_failed = true;
        if (inf[13] != 0.0) break;                   // capacity reached (RbfModel.jl:402) or no candidates left
        // ---- 3. rank-4 update
        if (has_tile && tK > K) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double4 i4 = ld4(cAs + q * LD, tI, H), k4 = ld4(cA + q * LD, tK, H);
                const double ai[4] = {i4.x, i4.y, i4.z, i4.w}, ak[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int c = 0; c < 4; ++c) A[a][c] = fma(-ai[a], ak[c], A[a][c]);
            }
        }
        if (K < 25 && has_tile && tK == K + 1 && tI == K + 1) SCHUR_STAMPX(128 + 8 * K + 6);
    }
    SCHUR_STAMP(5);
    // thread 0 owns the last diagonal tile: it is alive until the end and has seen every accepted pivot
    if (nacc_diag >= 0) nacc = nacc_diag;
    if (tid == 0 && bm && (bm_failed || nacc != mc)) { P.n_r4[b] = -1; return; }      // build mode: the general kernel takes the instance
    if (tid == 0) {
        P.n_r4[b] = npre + nacc; if (P.status) P.status[b] = 0;
        if (bm && P.alpha2_out) P.alpha2_out[b] = rf.alpha2;
        if (keep) {
            keep[g.off_acc + MC + 0] = inv_s; keep[g.off_acc + MC + 1] = (double)N0; keep[g.off_acc + MC + 2] = (double)nacc;
            keep[g.off_acc + MC + 3] = (double)mc;
            P.elig[b] = 1;
        }
    }
}

template <bool SMALL>
static cudaError_t launch_elim(const Round4Params& P, const SchurGeom& g, int nthreads, size_t smem, int rpl, cudaStream_t s) {
    cudaError_t e;
#define MRBF_ELIM_CASE(R)                                                                                                          \
    case R:                                                                                                                        \
        if (P.dbg_clock) {                                                                                                         \
            e = raise_dyn_smem(round4_elim_kernel<SMALL, R, true>, smem);  \
            if (e != cudaSuccess) return e;                                                                                        \
            round4_elim_kernel<SMALL, R, true><<<P.B, nthreads, smem, s>>>(P, g);                                                  \
        } else {                                                                                                                   \
            e = raise_dyn_smem(round4_elim_kernel<SMALL, R, false>, smem); \
            if (e != cudaSuccess) return e;                                                                                        \
            round4_elim_kernel<SMALL, R, false><<<P.B, nthreads, smem, s>>>(P, g);                                                 \
        }                                                                                                                          \
        break;
    switch (rpl) {
        MRBF_ELIM_CASE(1)
        MRBF_ELIM_CASE(2)
        MRBF_ELIM_CASE(3)
    default: return cudaErrorInvalidValue;
    }
#undef MRBF_ELIM_CASE
    return cudaGetLastError();
}

cudaError_t launch_round4_schur(const Round4Params& P, const SchurGeom& g, cudaStream_t s) {
    const int p = poly_dim(P.n, P.cfg.polynomial_degree);
    const int rpl = (p + 31) / 32;
    const size_t psmem = g.ps_doubles * sizeof(double);
    cudaError_t e = raise_dyn_smem(round4_prep_kernel, psmem);
    if (e != cudaSuccess) return e;
    round4_prep_kernel<<<P.B, 256, psmem, s>>>(P, g);
    if (g.two_variants) {
        e = launch_elim<true>(P, g, 384, g.smem_small * sizeof(double), rpl, s);
        if (e == cudaSuccess) e = launch_elim<false>(P, g, g.nthreads, g.smem_doubles * sizeof(double), rpl, s);
    } else if (g.nthreads <= 384) {
        e = launch_elim<true>(P, g, g.nthreads, g.smem_doubles * sizeof(double), rpl, s);
    } else {
        e = launch_elim<false>(P, g, g.nthreads, g.smem_doubles * sizeof(double), rpl, s);
    }
    return e;
}

}  // namespace mrbf
