// Training-set search kernels (rounds 1-4 of prepare_update_model), one CTA per instance.
//
// Reference behaviour restated (never copied): src/models/RbfModel.jl:205-307, 352-499, 518-655;
// src/models/AffinelyIndependentPoints.jl:4-106; src/Databases.jl:324-327; src/utilities.jl:126-221, 437-448.
//
// B200 design notes
//   * One CTA per instance; instances are independent, so the grid is the batch (C3: 4096 CTAs).
//   * The greedy filter keeps the trailing block W of the Householder Q *explicitly* and updates it by
//     one reflector per accepted point (O(n (n-j)) instead of the reference's full re-factorisation,
//     O(n j^2)); same reflectors as LAPACK geqrf (beta = -sign(alpha) ||x||), so Z matches to rounding.
//   * Round 4 never forms the dense Givens matrix G or the dense (N+1)^3 product of RbfModel.jl:462: only the
//     last row of G is needed (closed form from the c_j, s_j), Q is kept as its first p columns, and the
//     rotations are applied to column pairs.  Per candidate: O(N^2) coalesced mat-vecs instead of O(N^3).
//   * State lives in shared memory when it fits (small n / few points), else in a per-instance global
//     workspace that stays L2-resident; all mat-vecs are laid out so that consecutive threads read
//     consecutive addresses.
#include "mrbf_common.cuh"
#include "mrbf_kernels.h"

namespace mrbf {

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool in_box(const double* s, const double* lb, const double* ub, int n) {
    bool ok = true;
    for (int i = 0; i < n; ++i) ok = ok && (lb[i] <= s[i]) && (s[i] <= ub[i]);
    return ok;
}

// ------------------------------------------------------------------------------------------------
// Rounds 1-3
// ------------------------------------------------------------------------------------------------
struct FilterState {
    int n, ldz;
    double* W;        // n x (n - jY) trailing block of Q (unnormalised), column-major, ld = ldz
    double* Z;        // same, columns scaled by their inf-norm
    double* xp;       // n   scratch: W' y
    double* u;        // n   scratch: W v
    double* vv;       // n   scratch: reflector
    double* red;      // 80 doubles reduction scratch
    int* redi;        // 40 ints
    int* cl;          // db_stride ints: compacted candidate list of the current run
    long long* dbg;   // instrumentation: clock stamps of CTA 0 (or NULL)
    int jY;           // columns of Y so far
};

// cflag bits
#define CF_BOX1 1
#define CF_BOX2 2
#define CF_USED 4

// Block arg-max over non-negative scores with "first maximiser" tie-breaking (smallest position wins), one barrier: the warp stage
// compares the scores as integers with redux.sync (IEEE doubles >= 0 order like their bit patterns), the eight warp winners are
// combined by every thread.  `par` alternates the scratch slots so that consecutive calls need no protective barrier.
__device__ __forceinline__ ArgMax block_argmax_pos(ArgMax m, double* redv, int* redi, int par) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    const double v = (m.id >= 0) ? m.v : 0.0;
    const unsigned vh = (unsigned)__double2hiint(v), vl = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_max_sync(0xffffffffu, vh);
    const unsigned ml = __reduce_max_sync(0xffffffffu, vh == mh ? vl : 0u);
    const unsigned mid = __reduce_min_sync(0xffffffffu, (vh == mh && vl == ml && m.id >= 0) ? (unsigned)m.id : 0x7fffffffu);
    if (lane == 0) { redv[par * 16 + warp] = __hiloint2double((int)mh, (int)ml); redi[par * 16 + warp] = (mid == 0x7fffffffu) ? -1 : (int)mid; }
    __syncthreads();
    ArgMax r; r.v = 0.0; r.id = -1;
    for (int w = 0; w < nw; ++w) { ArgMax c_; c_.v = redv[par * 16 + w]; c_.id = redi[par * 16 + w]; r = better(r, c_); }
    return r;
}

// One run of the affinely-independent filter over the candidates with (cflags & want) == want and
// !(cflags & (CF_USED | avoid)).  Picks are appended to out[]; returns their number.
//
// The candidates of the run are compacted into a dense list cl[] (ascending ids, so "first maximiser" = smallest position).
// T holds, per list position, the projection coefficients y = W' s on the current trailing block W (coordinate-major).
// They are updated with the same reflector that updates W (O(n-j) per candidate and step) instead of being recomputed
// (O(n (n-j))), and the scores || Z (Z' s) ||_inf = || (W D^-2) y ||_inf  (D = column inf-norms of W, AffinelyIndependentPoints.jl:8)
// of all candidates are one small GEMM per step, register-tiled 4 rows x 4 candidates per thread.
__device__ __forceinline__ int filter_run(FilterState& st, const double* S /* shifted seeds, coordinate-major: S[i*ldS + id] */, double* T,
                          int ldS, unsigned char* cflags, int n_db, unsigned want, unsigned avoid, double piv, int n_wanted, int* out) {
    const int n = st.n, ldz = st.ldz, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    int* cl = st.cl;
    double* Wt = st.Z;                            // W D^-2 while the run lasts; the real Z is written at the end
    int found = 0;
    // ---- compact candidate list (one warp; n_db is small)
    if (warp == 0) {
        int base = 0;
        for (int i0 = 0; i0 < n_db; i0 += 32) {
            const int id = i0 + lane;
            bool act = false;
            if (id < n_db) { const unsigned f = cflags[id]; act = ((f & want) == want) && !(f & (CF_USED | avoid)); }
            const unsigned msk = __ballot_sync(0xffffffffu, act);
            if (act) cl[base + __popc(msk & ((1u << lane) - 1u))] = id;
            base += __popc(msk);
        }
        if (lane == 0) st.redi[36] = base;
    }
    __syncthreads();
    const int nc = st.redi[36];
    if (nc == 0) return 0;
    // ---- y = W' s for every candidate of this run, and the first pick: argmax ||s||_inf, first maximiser, unconditional
    ArgMax mine; mine.v = 0.0; mine.id = -1;
    {
        const int nw0 = n - st.jY;
        const int G = (nc > 128) ? 1 : ((nc > 64) ? 2 : ((nc > 32) ? 4 : 8));
        const int cpp = nt / G;
        for (int c0 = 0; c0 < nc; c0 += cpp) {
            const int pos = c0 + tid / G, h = tid % G;
            if (pos >= nc) continue;
            const double* s = S + cl[pos];
            for (int c = h; c < nw0; c += G) {
                const double* wc = st.W + c * ldz;
                double a0 = 0.0, a1 = 0.0;
                int i = 0;
                for (; i + 1 < n; i += 2) { a0 = fma(wc[i], s[i * ldS], a0); a1 = fma(wc[i + 1], s[(i + 1) * ldS], a1); }
                if (i < n) a0 = fma(wc[i], s[i * ldS], a0);
                T[c * ldS + pos] = a0 + a1;
            }
            if (h == 0) {
                double v = 0.0;
                for (int i = 0; i < n; ++i) v = fmax(v, fabs(s[i * ldS]));
                ArgMax c_; c_.v = v; c_.id = pos;
                mine = better(mine, c_);
            }
        }
    }
    int apar = 0;
    ArgMax best = block_argmax_pos(mine, st.red, st.redi, apar); apar ^= 1;
    const int RQ = (n + 3) >> 2;                  // row quads of the scoring GEMM
    int RQ2 = 1; while (RQ2 < RQ && RQ2 < 32) RQ2 <<= 1;
#define FSTAMP(i) do { if (st.dbg && threadIdx.x == 0 && found < 4) st.dbg[8 + 8 * found + (i)] = clock64(); } while (0)
    for (;;) {
        FSTAMP(0);
        // ---- accept position best.id: Y <- [Y s]; its projection on W is already in T.  dlarfg on that column, one warp.
        const int nw = n - st.jY;                 // columns of W before the update
        const int bpos = best.id;
        // Every thread derives the reflector of the picked column for itself (broadcast reads of that column of T, ~100 instructions) --
        // cheaper than one warp doing it while seven wait at a barrier.  The picked candidate is excluded from the update below, so its
        // column stays intact until the barrier at the end of the phase.
        //   beta = -sign(alpha) ||(alpha, x)||, tau = (beta - alpha) / beta, v = [1; x(2:) / (alpha - beta)]   (LAPACK dlarfg; the values are
        //   O(Delta), so no rescaling loop is needed; reciprocal-root / reciprocal seeds + Newton instead of hypot and two divisions)
        const double* xcol = T + bpos;
        double tau = 0.0, sc = 0.0;
        {
            double s0 = 0.0, s1 = 0.0;
            int c = 1;
            for (; c + 1 < nw; c += 2) { const double x0 = xcol[c * ldS], x1 = xcol[(c + 1) * ldS]; s0 = fma(x0, x0, s0); s1 = fma(x1, x1, s1); }
            if (c < nw) { const double x0 = xcol[c * ldS]; s0 = fma(x0, x0, s0); }
            const double sig = s0 + s1, alpha = xcol[0];
            if (sig != 0.0) {
                const double nn = fma(alpha, alpha, sig);
                const double beta = -copysign(nn * fast_rsqrt(nn), alpha);
                tau = (beta - alpha) * fast_rcp(beta);
                sc = fast_rcp(alpha - beta);
            }
        }
        if (tid == 0) {
            const int id = cl[bpos];
            cflags[id] |= CF_USED;
            out[found] = id + 1;
        }
        FSTAMP(1);
        // ---- the reflector on W (row-private) and on the candidates' coefficients (position-private).  When the block has two
        // threads per task, the column range of a task is split between a lane pair (half the serial chain).
        {
            const int ntask = n + nc;
            const int SP = (2 * ntask <= nt) ? 2 : 1;
            const int half = (SP == 2) ? (tid & 1) : 0;
            const int mid = (SP == 2) ? ((nw + 1) >> 1) : nw;        // half 0: columns [0, mid), half 1: [mid, nw)
            const int c_lo = half ? mid : 0, c_hi = half ? nw : mid;
            for (int w0 = 0; w0 < ntask; w0 += nt / SP) {
                const int w = w0 + tid / SP;
                const bool act = w < ntask;
                double* base_ = nullptr; int ldx = 0;
                if (act) { if (w < n) { base_ = st.W + w; ldx = ldz; } else { base_ = T + (w - n); ldx = ldS; } }
                const bool upd_ = act && (w - n != bpos);      // the picked column is read by everybody in this phase: leave it alone
                double a0 = 0.0, a1 = 0.0;
                if (upd_) {
                    int c = c_lo;
                    if (c == 0 && c < c_hi) { a0 = base_[0]; c = 1; }                                   // v_0 = 1
                    for (; c + 1 < c_hi; c += 2) { a0 = fma(base_[c * ldx], xcol[c * ldS] * sc, a0); a1 = fma(base_[(c + 1) * ldx], xcol[(c + 1) * ldS] * sc, a1); }
                    if (c < c_hi) a0 = fma(base_[c * ldx], xcol[c * ldS] * sc, a0);
                }
                double a = a0 + a1;
                if (SP == 2) a += __shfl_xor_sync(0xffffffffu, a, 1);
                a *= tau;
                // x'[c-1] = x[c] - a v_c for c = 1..nw-1: half 1 overwrites position mid-1, which half 0 still needs as a source
                double keep_ = 0.0;
                if (upd_ && SP == 2 && half == 0 && mid - 1 >= 1) keep_ = base_[(mid - 1) * ldx];
                if (SP == 2) __syncwarp();
                if (upd_) {
                    const int lo = (c_lo < 1) ? 1 : c_lo;
                    for (int c = lo; c < c_hi; ++c) {
                        const double src = (SP == 2 && half == 0 && c == mid - 1) ? keep_ : base_[c * ldx];
                        base_[(c - 1) * ldx] = fma(-a * sc, xcol[c * ldS], src);
                    }
                }
            }
        }
        __syncthreads();
        FSTAMP(2);
        if (tid == 0) cl[bpos] = ~cl[bpos];      // stays in the list (its column of T is dead) but never wins again
        st.jY += 1;
        found += 1;
        const int zc = n - st.jY;
        if (found == n_wanted) break;
        // ---- Wt = W D^-2, D = column inf-norms of W (AffinelyIndependentPoints.jl:8): one warp per column
        for (int c = warp; c < zc; c += nwarps) {
            const double* wc = st.W + c * ldz;
            double mx = 0.0;
            for (int i = lane; i < n; i += 32) mx = fmax(mx, fabs(wc[i]));
            {                                             // warp max of non-negative doubles as integers (two redux instead of five shuffles)
                const unsigned vh = (unsigned)__double2hiint(mx), vl = (unsigned)__double2loint(mx);
                const unsigned mh = __reduce_max_sync(0xffffffffu, vh);
                const unsigned ml = __reduce_max_sync(0xffffffffu, vh == mh ? vl : 0u);
                mx = __hiloint2double((int)mh, (int)ml);
            }
            const double r = fast_rcp(mx * mx);
            for (int i = lane; i < n; i += 32) Wt[c * ldz + i] = wc[i] * r;
        }
        __syncthreads();
        if (st.dbg && threadIdx.x == 0 && found <= 4) st.dbg[8 + 8 * (found - 1) + 3] = clock64();
        // ---- scores: thread = (row quad rq, candidate quad cq); acc[4 rows][4 candidates] += Wt[rows, q] * y[q, candidates]
        mine.v = 0.0; mine.id = -1;
        {
            const int CQ = (nc + 3) >> 2;
            const int rq0 = tid % RQ2, cqs = nt / RQ2;
            for (int cq = tid / RQ2; cq < ((CQ + cqs - 1) / cqs) * cqs; cq += cqs) {
                double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;       // max |.| over this thread's rows, per candidate of the quad
                if (cq < CQ)
                    for (int rq = rq0; rq < RQ; rq += RQ2) {
                        const int i0 = 4 * rq;
                        double acc[4][4];
#pragma unroll
                        for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
                            for (int c_ = 0; c_ < 4; ++c_) acc[a_][c_] = 0.0;
                        const double* wp = Wt + i0; const double* yp = T + 4 * cq;
                        for (int q = 0; q < zc; ++q) {
                            const double2 wa = *reinterpret_cast<const double2*>(wp + q * ldz), wb = *reinterpret_cast<const double2*>(wp + q * ldz + 2);
                            const double w0 = wa.x, w1 = wa.y, w2 = wb.x, w3 = wb.y;      // rows >= n of Wt are zero
                            const double2 ya = *reinterpret_cast<const double2*>(yp + q * ldS), yb = *reinterpret_cast<const double2*>(yp + q * ldS + 2);
                            acc[0][0] = fma(w0, ya.x, acc[0][0]); acc[0][1] = fma(w0, ya.y, acc[0][1]); acc[0][2] = fma(w0, yb.x, acc[0][2]); acc[0][3] = fma(w0, yb.y, acc[0][3]);
                            acc[1][0] = fma(w1, ya.x, acc[1][0]); acc[1][1] = fma(w1, ya.y, acc[1][1]); acc[1][2] = fma(w1, yb.x, acc[1][2]); acc[1][3] = fma(w1, yb.y, acc[1][3]);
                            acc[2][0] = fma(w2, ya.x, acc[2][0]); acc[2][1] = fma(w2, ya.y, acc[2][1]); acc[2][2] = fma(w2, yb.x, acc[2][2]); acc[2][3] = fma(w2, yb.y, acc[2][3]);
                            acc[3][0] = fma(w3, ya.x, acc[3][0]); acc[3][1] = fma(w3, ya.y, acc[3][1]); acc[3][2] = fma(w3, yb.x, acc[3][2]); acc[3][3] = fma(w3, yb.y, acc[3][3]);
                        }
#pragma unroll
                        for (int a_ = 0; a_ < 4; ++a_) {
                            v0 = fmax(v0, fabs(acc[a_][0])); v1 = fmax(v1, fabs(acc[a_][1])); v2 = fmax(v2, fabs(acc[a_][2])); v3 = fmax(v3, fabs(acc[a_][3]));
                        }
                    }
                for (int o = RQ2 >> 1; o > 0; o >>= 1) {
                    v0 = fmax(v0, __shfl_xor_sync(0xffffffffu, v0, o)); v1 = fmax(v1, __shfl_xor_sync(0xffffffffu, v1, o));
                    v2 = fmax(v2, __shfl_xor_sync(0xffffffffu, v2, o)); v3 = fmax(v3, __shfl_xor_sync(0xffffffffu, v3, o));
                }
                if (rq0 == 0 && cq < CQ) {
                    const double vs[4] = {v0, v1, v2, v3};
#pragma unroll
                    for (int c_ = 0; c_ < 4; ++c_) {
                        const int pos = 4 * cq + c_;
                        if (pos < nc && cl[pos] >= 0) { ArgMax cnd; cnd.v = vs[c_]; cnd.id = pos; mine = better(mine, cnd); }
                    }
                }
            }
        }
        if (st.dbg && threadIdx.x == 0 && found <= 4) st.dbg[8 + 8 * (found - 1) + 4] = clock64();
        best = block_argmax_pos(mine, st.red, st.redi, apar); apar ^= 1;
        if (st.dbg && threadIdx.x == 0 && found <= 4) st.dbg[8 + 8 * (found - 1) + 5] = clock64();
        if (best.id < 0) break;                   // no more candidates
        if (!(best.v > piv)) break;               // AffinelyIndependentPoints.jl:92
    }
    {                                             // Z = W ./ colmax|W| for the caller (improving directions)
        const int zc = n - st.jY;
        __syncthreads();
        for (int c = warp; c < zc; c += nwarps) {
            const double* wc = st.W + c * ldz;
            double mx = 0.0;
            for (int i = lane; i < n; i += 32) mx = fmax(mx, fabs(wc[i]));
            mx = warp_max(mx);
            for (int i = lane; i < n; i += 32) st.Z[c * ldz + i] = wc[i] / mx;
        }
        __syncthreads();
    }
    return found;
}

template <bool WZS, bool STS, int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? 3 : 1) select_rounds123_kernel(SelectParams P) {
    extern __shared__ double smem[];
    const int b = blockIdx.x, n = P.n, tid = threadIdx.x, nt = blockDim.x;
    const int ldz = (n + 3) & ~3;              // rows padded to a multiple of 4 (zero rows): the scoring tiles load row quads
    double* x = smem;
    double* lb1 = x + n; double* ub1 = lb1 + n; double* lb2 = ub1 + n; double* ub2 = lb2 + n;
    double* xp = ub2 + n; double* u = xp + n; double* vv = u + n;
    double* red = vv + n;                        // 80
    int* redi = (int*)(red + 80);                // 40 ints = 20 doubles
    int* ctl = redi + 40;                        // 8 ints control
    int* cl = ctl + 8;                           // db_stride ints (compacted candidate list)
    double* wz = red + 80 + 24 + 2 * ((P.db_stride + 3) / 4);
    double* W; double* S; double* T;
    const int ldS = (P.db_stride + 1) & ~1;      // even: the scoring tiles read coefficient pairs
    if constexpr (WZS) W = wz; else W = P.WZ + (size_t)b * 2 * n * ldz;
    double* Z = W + n * ldz;
    // shifted seeds (read once per run) stay in the L2-resident global workspace; the projection coefficients T, which every
    // step reads and updates, live in shared memory when they fit
    S = P.S + (size_t)b * ldS * n;
    if constexpr (STS) T = wz + 2 * n * ldz; else T = P.T + (size_t)b * ldS * n;
    const int n_db = P.n_db[b];
    const double* sites = P.sites + (size_t)b * P.db_stride * n;
    unsigned char* cflags = P.cflags + (size_t)b * P.db_stride;
    const int x_index = P.x_index[b] - 1;
    const double delta = P.delta[b];
    const double delta_1 = P.cfg.theta_enlarge_1 * delta;
    const double piv = P.cfg.theta_pivot * delta_1;
    const double delta_2 = P.cfg.theta_enlarge_2 * P.delta_max;
    int* r1 = P.r1 + (size_t)b * n; int* r2 = P.r2 + (size_t)b * n;
    double* r3s = P.r3_sites + (size_t)b * n * n;
    double* dirs = P.dirs + (size_t)b * n * n;

    for (int i = tid; i < n; i += nt) {
        double xi = P.x[(size_t)b * n + i];
        x[i] = xi;
        lb1[i] = fmax(P.glb[i], xi - delta_1); ub1[i] = fmin(P.gub[i], xi + delta_1);   // utilities.jl:290-294
        lb2[i] = fmax(P.glb[i], xi - delta_2); ub2[i] = fmin(P.gub[i], xi + delta_2);
    }
    __syncthreads();
    // box scan (Databases.jl:324-327) + shifted seeds; also writes the box-2 bounds for round 4
    for (int id = tid; id < n_db; id += nt) {
        const double* s = sites + (size_t)id * n;
        unsigned f = 0;
        if (id != x_index) {
            if (in_box(s, lb1, ub1, n)) f |= CF_BOX1;
            if (in_box(s, lb2, ub2, n)) f |= CF_BOX2;
        }
        cflags[id] = (unsigned char)f;
        for (int i = 0; i < n; ++i) S[i * ldS + id] = s[i] - x[i];
    }
    for (int i = tid; i < n; i += nt) { P.lb2[(size_t)b * n + i] = lb2[i]; P.ub2[(size_t)b * n + i] = ub2[i]; }

    bool ensure_fl = P.flags_in[2 * b] != 0;
    bool force_rebuild = P.flags_in[2 * b + 1] != 0;
    bool rebuilt = false;
    FilterState st; st.n = n; st.ldz = ldz; st.W = W; st.Z = Z; st.xp = xp; st.u = u; st.vv = vv; st.red = red; st.redi = redi; st.cl = cl; st.dbg = (b == 0) ? P.dbg_clock : nullptr;
    if (st.dbg && tid == 0) st.dbg[0] = clock64();
    int n_r1, n_r2, n_r3, n_dirs;
    bool fully_linear;
    for (;;) {   // at most two passes: the second is the coordinate rebuild (RbfModel.jl:634-637)
        __syncthreads();
        for (int e = tid; e < ldz * n; e += nt) { int i = e % ldz, c = e / ldz; W[i + c * ldz] = (i == c) ? 1.0 : 0.0; Z[i + c * ldz] = (i == c) ? 1.0 : 0.0; }
        for (int id = tid; id < n_db; id += nt) cflags[id] &= (unsigned char)~CF_USED;
        __syncthreads();
        st.jY = 0; n_r1 = n_r2 = n_r3 = 0; fully_linear = false;
        const bool skip_search = force_rebuild || !P.cfg.optimized_sampling;
        if (skip_search) {                       // RbfModel.jl:564-569: directions e_1..e_n in natural order
            for (int e = tid; e < n * n; e += nt) dirs[e] = ((e % n) == (e / n)) ? 1.0 : 0.0;
            n_dirs = n;
        } else {
            n_r1 = filter_run(st, S, T, ldS, cflags, n_db, CF_BOX1, 0, piv, n, r1);
            const int zc = n - st.jY;            // improving directions = reverse(eachcol(Z)), RbfModel.jl:232
            for (int e = tid; e < n * zc; e += nt) { int i = e % n, c = e / n; dirs[i + (size_t)c * n] = Z[i + (zc - 1 - c) * ldz]; }
            n_dirs = zc;
        }
        int n_missing = n - n_r1;
        const bool approx = fabs(delta - P.delta_max) <= P.approx_rtol * fmax(fabs(delta), fabs(P.delta_max));
        if (n_missing == 0 || skip_search || ensure_fl || (approx && P.cfg.theta_enlarge_1 == P.cfg.theta_enlarge_2)) {
            fully_linear = true;                 // RbfModel.jl:588-591
        } else {                                 // round 2: box 2, excluding every round-1 candidate
            n_r2 = filter_run(st, S, T, ldS, cflags, n_db, CF_BOX2, CF_BOX1, piv, n_missing, r2);
        }
        n_missing -= n_r2;
        bool failed = false;
        if (n_missing > 0) {                     // round 3, RbfModel.jl:269-307
            int n_new = min(n_missing, P.max_new[b]); if (n_new < 0) n_new = 0;
            bool fl = n_new >= n_missing;
            __syncthreads();
            if (tid == 0) ctl[0] = 0;
            __syncthreads();
            for (int i = tid; i < n_new; i += nt) {
                const double* d = dirs + (size_t)i * n;
                double len = intersect_box_absmax(n, x, d, lb1, ub1);
                double on = 0.0;
                for (int r = 0; r < n; ++r) { double o = len * d[r]; r3s[(size_t)i * n + r] = x[r] + o; on = fmax(on, fabs(o)); }
                if (on <= piv) atomicOr(&ctl[0], 1);
            }
            __syncthreads();
            const bool any_small = ctl[0] != 0;
            if (any_small) {
                if (ensure_fl && !force_rebuild) failed = true;
                else fl = false;
            }
            if (!failed) { n_r3 = n_new; fully_linear = fl && (n_r2 == 0); }
        }
        if (!failed) break;
        force_rebuild = true; ensure_fl = true; rebuilt = true;
    }
    __syncthreads();
    if (st.dbg && tid == 0) st.dbg[1] = clock64();
    if (tid == 0) {
        P.n_r1[b] = n_r1; P.n_r2[b] = n_r2; P.n_r3[b] = n_r3; P.n_dirs[b] = n_dirs;
        P.flags_out[2 * b] = fully_linear ? 1 : 0; P.flags_out[2 * b + 1] = rebuilt ? 1 : 0;
        // found set for round 4: [centre; r1; r2] as ids, round-3 sites as extra sites
        int* found = P.found + (size_t)b * P.found_stride;
        int nf = 0;
        found[nf++] = x_index + 1;
        for (int i = 0; i < n_r1; ++i) found[nf++] = r1[i];
        for (int i = 0; i < n_r2; ++i) found[nf++] = r2[i];
        P.n_found[b] = nf;
    }
}

// ------------------------------------------------------------------------------------------------
// Round 4 (Wild's bounded-Cholesky augmentation), RbfModel.jl:352-499
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) round4_kernel(Round4Params P) {
    extern __shared__ double smem[];
    const int b = P.b0 + blockIdx.x, n = P.n, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    if (P.only_marked && P.n_r4[b] != -1) return;      // handled by the shared-memory fast path
    const bool prefix = P.prefix_mode != 0;            // walk under-poised instances until N = p only (see Round4Params::hyb)
    if (prefix) {
        const int N0_ = P.n_found[b] + (P.n_extra ? P.n_extra[b] : 0);
        if (tid == 0) { P.hyb[b] = 0; P.pre_cnt[b] = 0; P.pre_min[b] = 0; }
        if (!(N0_ < poly_dim(P.n, P.cfg.polynomial_degree)) || !(N0_ < P.max_points) || N0_ > P.NM) return;     // regular instance
    }
    const int NM = P.NM, MM = P.NM;
    const int deg = P.cfg.polynomial_degree;
    const int p = poly_dim(n, deg);
    const int pl = p > 0 ? p : 1;
    // shared vectors
    double* xi = smem;                 // n
    double* phix = xi + n;             // NM
    double* q = phix + NM;             // NM
    double* u = q + NM;                // NM
    double* v = u + NM;                // NM
    double* t = v + NM;                // NM
    double* cs = t + NM;               // pl
    double* sn = cs + pl;              // pl
    double* gt = sn + pl;              // pl
    double* rl = gt + pl;              // pl
    double* red = rl + pl;             // 80
    double* mats = red + 80;
    double* ws = P.ws_in_smem ? mats : P.ws + (size_t)blockIdx.x * P.ws_stride;
    double* Ct = ws;                               // NM x n   (Ct[k*NM + i] = centre i, coordinate k)
    double* Phi = Ct + (size_t)NM * n;             // NM x NM
    double* Q1 = Phi + (size_t)NM * NM;            // NM x pl
    double* R = Q1 + (size_t)NM * pl;              // pl x pl
    double* Zt = R + (size_t)pl * pl;              // MM x NM  (Zt[c + i*MM] = Z[i, c])
    double* Li = Zt + (size_t)MM * NM;             // MM x MM  lower triangular inverse Cholesky factor

    const int n_db = P.n_db[b];
    const double* sites = P.sites + (size_t)b * P.db_stride * n;
    const double* lb2 = P.lb2 + (size_t)b * n;
    const double* ub2 = P.ub2 + (size_t)b * n;
    const int* found = P.found + (size_t)b * P.found_stride;
    const int nf_ids = P.n_found[b];
    const int n_extra = P.n_extra ? P.n_extra[b] : 0;
    const double* extra = P.extra_sites ? P.extra_sites + (size_t)b * P.extra_stride * n : nullptr;
    int* r4 = P.r4 + (size_t)b * P.r4_stride;
    int N = nf_ids + n_extra;
    const int max_points = P.max_points;
    int nr4 = 0, m = 0;
    if (!(N < max_points) || N > NM) { if (tid == 0) { P.n_r4[b] = 0; if (P.status) P.status[b] = (N > NM) ? -1 : 0; } return; }
    // candidates (results_in_box_indices(db, lb_2, ub_2, indices_found_so_far), RbfModel.jl:360), ascending ids
    unsigned char* cand = P.cand + (size_t)b * P.db_stride;
    for (int id = tid; id < n_db; id += nt) {
        bool ok = in_box(sites + (size_t)id * n, lb2, ub2, n);
        for (int f = 0; f < nf_ids && ok; ++f) ok = (found[f] != id + 1);
        cand[id] = ok ? 1 : 0;
    }

    // ---- centres, Phi, Pi -> Householder QR -> Q1 (first p columns of the full Q), R
    for (int e = tid; e < N * n; e += nt) {
        int i = e / n, k = e % n;
        double val = (i < nf_ids) ? sites[(size_t)(found[i] - 1) * n + k] : extra[(size_t)(i - nf_ids) * n + k];
        Ct[(size_t)k * NM + i] = val;
    }
    __syncthreads();
    for (int e = tid; e < N * N; e += nt) {
        int i = e % N, j = e / N;
        double r2 = 0.0;
        for (int k = 0; k < n; ++k) { double d = Ct[(size_t)k * NM + i] - Ct[(size_t)k * NM + j]; r2 = fma(d, d, r2); }
        Phi[i + (size_t)j * NM] = rad_phi(P.rf, r2);
    }
    // Pi into Q1's storage (N x p), factor in place, keep reflectors in Q1 below the diagonal temporarily
    for (int e = tid; e < N * p; e += nt) {
        int i = e % N, c = e / N;
        Q1[i + (size_t)c * NM] = (c == 0) ? 1.0 : Ct[(size_t)(c - 1) * NM + i];
    }
    for (int e = tid; e < pl * pl; e += nt) R[e] = 0.0;
    __syncthreads();
    const int kq = N < p ? N : p;                  // number of reflectors
    // dgeqr2 on Q1 (N x p); taus kept in `t` (shared)
    for (int j = 0; j < kq; ++j) {
        double part = 0.0;
        for (int i = j + 1 + tid; i < N; i += nt) { double a = Q1[i + (size_t)j * NM]; part = fma(a, a, part); }
        double xn2 = block_sum(part, red);
        if (tid == 0) {
            double alpha = Q1[j + (size_t)j * NM], xnorm = sqrt(xn2), tau = 0.0, sc = 0.0, beta = alpha;
            if (xnorm != 0.0 && j + 1 < N) {
                beta = -copysign(hypot(alpha, xnorm), alpha);
                tau = (beta - alpha) / beta; sc = 1.0 / (alpha - beta);
            }
            t[j] = tau; red[70] = sc; red[71] = beta;
        }
        __syncthreads();
        const double tau = t[j], sc = red[70];
        for (int i = j + 1 + tid; i < N; i += nt) Q1[i + (size_t)j * NM] *= sc;
        if (tid == 0) Q1[j + (size_t)j * NM] = red[71];
        __syncthreads();
        if (tau != 0.0)
            for (int c = j + 1 + warp; c < p; c += nwarps) {      // warp per trailing column
                double* col = Q1 + (size_t)c * NM; const double* vj = Q1 + (size_t)j * NM;
                double a = 0.0;
                for (int i = j + 1 + lane; i < N; i += 32) a = fma(vj[i], col[i], a);
                a = warp_sum(a) + col[j];
                a *= tau;
                for (int i = j + 1 + lane; i < N; i += 32) col[i] = fma(-a, vj[i], col[i]);
                __syncwarp();
                if (lane == 0) col[j] -= a;
            }
        __syncthreads();
    }
    // R (upper trapezoidal, kq x p) out of Q1's upper part
    for (int e = tid; e < kq * p; e += nt) { int r = e % kq, c = e / kq; if (r <= c) R[r + (size_t)c * pl] = Q1[r + (size_t)c * NM]; }
    __syncthreads();
    // explicit first min(N,p) columns of Q: apply H_0..H_{kq-1} (in reverse) to e_c; reflector j lives in column j.
    // Done column by column into `Zt` scratch (N x kq), then copied over Q1.
    {
        double* Qx = Zt;                           // scratch, ld = NM (Zt is MM*NM >= NM*pl doubles)
        for (int c = warp; c < kq; c += nwarps) {
            double* col = Qx + (size_t)c * NM;
            for (int i = lane; i < N; i += 32) col[i] = (i == c) ? 1.0 : 0.0;
            __syncwarp();
            for (int j = kq - 1; j >= 0; --j) {
                const double tau = t[j];
                if (tau == 0.0) continue;
                const double* vj = Q1 + (size_t)j * NM;
                double a = 0.0;
                for (int i = j + 1 + lane; i < N; i += 32) a = fma(vj[i], col[i], a);
                a = warp_sum(a) + col[j];
                a *= tau;
                for (int i = j + 1 + lane; i < N; i += 32) col[i] = fma(-a, vj[i], col[i]);
                __syncwarp();
                if (lane == 0) col[j] -= a;
                __syncwarp();
            }
        }
        __syncthreads();
        for (int e = tid; e < N * kq; e += nt) { int i = e % N, c = e / N; Q1[i + (size_t)c * NM] = Qx[i + (size_t)c * NM]; }
        // when N < p the remaining columns of Q (kq..N-1) do not exist (Q is N x N = N x kq)
        __syncthreads();
    }
    const double phi0 = Phi[0];
    const int full_rank_dim = deg < 0 ? 0 : p;
    const double thr = P.chol_thr;                 // (theta_pivot_cholesky^2)^2, RbfModel.jl:370, 452

    int id_next = 0;
    for (int id = 0; id < n_db && N < max_points && nr4 < P.r4_stride && !(prefix && N >= p); ++id) {
        id_next = id + 1;
        if (!cand[id]) continue;                   // block-uniform: in box 2 and not in the found set
        __syncthreads();
        for (int k = tid; k < n; k += nt) xi[k] = sites[(size_t)id * n + k];
        __syncthreads();
        const int J = N < p ? N : p;
        if (warp == 0) {
            // Givens sequence on [R; pi_xi'] (utilities.jl:437-448), R untouched until acceptance
            for (int c = lane; c < p; c += 32) rl[c] = (c == 0) ? 1.0 : xi[c - 1];
            __syncwarp();
            for (int j = 0; j < J; ++j) {
                double c_, s_;
                givens(R[j + (size_t)j * pl], rl[j], c_, s_);
                __syncwarp();
                for (int c = j + lane; c < p; c += 32) rl[c] = -s_ * R[j + (size_t)c * pl] + c_ * rl[c];
                if (lane == 0) { cs[j] = c_; sn[j] = s_; }
                __syncwarp();
            }
            if (lane == 0) {
                double gh = 1.0;                   // last row of G: g_j = -s_j prod_{i>j} c_i, g_hat = prod c_i
                for (int j = J - 1; j >= 0; --j) { gt[j] = -sn[j] * gh; gh *= cs[j]; }
                red[72] = gh;
                double nr = 0.0;
                for (int c = 0; c < p; ++c) nr = hypot(nr, rl[c]);
                red[73] = nr;
            }
        } else {
            for (int i = tid - 32; i < N; i += nt - 32) {          // kernels(xi), RbfModel.jl:421
                double r2 = 0.0;
                for (int k = 0; k < n; ++k) { double d = xi[k] - Ct[(size_t)k * NM + i]; r2 = fma(d, d, r2); }
                phix[i] = rad_phi(P.rf, r2);
            }
        }
        __syncthreads();
        if (N < full_rank_dim && red[73] <= 2.220446049250313e-16 * 10) continue;   // RbfModel.jl:433-438
        const double gh = red[72];
        for (int i = tid; i < N; i += nt) {        // q = Q g~
            double a = 0.0;
            for (int j = 0; j < J; ++j) a = fma(Q1[i + (size_t)j * NM], gt[j], a);
            q[i] = a;
        }
        __syncthreads();
        double s1 = 0.0, s2 = 0.0;
        for (int i = tid; i < N; i += nt) {        // u = Phi q  (Phi symmetric, column-major => coalesced over i)
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int j = 0;
            for (; j + 4 <= N; j += 4) {
                a0 = fma(Phi[i + (size_t)j * NM], q[j], a0); a1 = fma(Phi[i + (size_t)(j + 1) * NM], q[j + 1], a1);
                a2 = fma(Phi[i + (size_t)(j + 2) * NM], q[j + 2], a2); a3 = fma(Phi[i + (size_t)(j + 3) * NM], q[j + 3], a3);
            }
            for (; j < N; ++j) a0 = fma(Phi[i + (size_t)j * NM], q[j], a0);
            double a = (a0 + a1) + (a2 + a3);
            s1 = fma(q[i], a, s1); s2 = fma(phix[i], q[i], s2);
            u[i] = fma(gh, phix[i], a);
        }
        block_sum2(s1, s2, red);                   // includes the barriers that publish u
        const double sigma = s1 + (2.0 * gh) * s2 + gh * gh * phi0;      // RbfModel.jl:447
        for (int c = tid; c < m; c += nt) {        // v = Z' u
            double a0 = 0.0, a1 = 0.0;
            int i = 0;
            for (; i + 2 <= N; i += 2) { a0 = fma(Zt[c + (size_t)i * MM], u[i], a0); a1 = fma(Zt[c + (size_t)(i + 1) * MM], u[i + 1], a1); }
            for (; i < N; ++i) a0 = fma(Zt[c + (size_t)i * MM], u[i], a0);
            v[c] = a0 + a1;
        }
        __syncthreads();
        double tn = 0.0;
        for (int r = tid; r < m; r += nt) {        // t = L^{-1} v
            double a0 = 0.0, a1 = 0.0;
            int c = 0;
            for (; c + 2 <= r + 1; c += 2) { a0 = fma(Li[r + (size_t)c * MM], v[c], a0); a1 = fma(Li[r + (size_t)(c + 1) * MM], v[c + 1], a1); }
            for (; c <= r; ++c) a0 = fma(Li[r + (size_t)c * MM], v[c], a0);
            double a = a0 + a1;
            t[r] = a; tn = fma(a, a, tn);
        }
        tn = block_sum(tn, red);
        const double nrm = sqrt(tn);
        const double tau2 = sigma - nrm * nrm;     // RbfModel.jl:449
        if (!(tau2 > thr)) continue;               // RbfModel.jl:452
        // ---- accept
        const double tv = sqrt(tau2);
        for (int i = tid; i <= N; i += nt) {       // Q <- blkdiag(Q,1) G': rotate column pairs (j, N)
            double bcol = (i == N) ? 1.0 : 0.0;
            for (int j = 0; j < J; ++j) {
                double a = (i == N) ? 0.0 : Q1[i + (size_t)j * NM];
                Q1[i + (size_t)j * NM] = cs[j] * a + sn[j] * bcol;
                bcol = -sn[j] * a + cs[j] * bcol;
            }
            if (N < p) Q1[i + (size_t)N * NM] = bcol;             // the new column becomes a pivot column next time
            else if (i == N) for (int j = J; j < p; ++j) Q1[i + (size_t)j * NM] = 0.0;
            // Z <- [Z q; 0 g_hat]
            Zt[m + (size_t)i * MM] = (i == N) ? gh : q[i];
        }
        for (int c = tid; c < m; c += nt) Zt[c + (size_t)N * MM] = 0.0;
        if (N < p) for (int j = J + tid; j < p; j += nt) if (j != N) Q1[N + (size_t)j * NM] = 0.0;
        // L^{-1} <- [L^{-1} 0; -(t' L^{-1})/tau 1/tau]   (warp per column, coalesced down the column)
        for (int c = warp; c < m; c += nwarps) {
            double a = 0.0;
            for (int r = c + lane; r < m; r += 32) a = fma(t[r], Li[r + (size_t)c * MM], a);
            a = warp_sum(a);
            if (lane == 0) v[c] = -a / tv;         // v is free now: holds the new row
        }
        if (warp == 0) {                           // R <- rotated [R; pi'] (only rows < p are non-zero)
            for (int c = lane; c < p; c += 32) rl[c] = (c == 0) ? 1.0 : xi[c - 1];
            __syncwarp();
            for (int j = 0; j < J; ++j) {
                for (int c = lane; c < p; c += 32) {
                    double a = R[j + (size_t)c * pl], bb = rl[c];
                    R[j + (size_t)c * pl] = cs[j] * a + sn[j] * bb;
                    rl[c] = -sn[j] * a + cs[j] * bb;
                }
                __syncwarp();
            }
            if (N < p) for (int c = lane; c < p; c += 32) R[N + (size_t)c * pl] = rl[c];
        }
        for (int i = tid; i < N; i += nt) { Phi[i + (size_t)N * NM] = phix[i]; Phi[N + (size_t)i * NM] = phix[i]; }
        for (int k = tid; k < n; k += nt) Ct[(size_t)k * NM + N] = xi[k];
        if (tid == 0) { Phi[N + (size_t)N * NM] = phi0; r4[nr4] = id + 1; }
        __syncthreads();
        for (int c = tid; c < m; c += nt) { Li[m + (size_t)c * MM] = v[c]; Li[c + (size_t)m * MM] = 0.0; }
        if (tid == 0) Li[m + (size_t)m * MM] = 1.0 / tv;
        N += 1; m += 1; nr4 += 1;
        __syncthreads();
    }
    if (tid == 0) {
        if (prefix && N >= p && id_next < n_db && N < max_points && nr4 < P.r4_stride) {
            P.hyb[b] = 1; P.pre_cnt[b] = nr4; P.pre_min[b] = id_next;       // poised now: the register kernels continue (n_r4 is theirs to write)
        } else {
            P.n_r4[b] = nr4; if (P.status) P.status[b] = 0;
            if (prefix) P.hyb[b] = 2;
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Round 4, Lagrange-basis formulation for the regular case N0 == p (found set = centre + n poised points): the blocked
// left-looking kernel for large databases (round4_elim_kernel in mrbf_round4_schur.cu takes databases of <= 128 sites).
//
// Same decisions as RbfModel.jl:420-452, different basis.  Let S0 be the found set (Pi_0 = Pi(S0) is p x p and
// non-singular).  Every later point xi has the null vector  n_xi = e_xi - sum_{s in S0} c_xi[s] e_s,
// c_xi = Pi_0^{-T} pi_xi, and these vectors span null(Pi') just like the reference's orthonormal Z.  With
// A = N' Phi N (the kernel matrix reduced to that basis) and its Cholesky pivot d^2 for xi,
//       tau^2 (reference, orthonormal basis)  =  g_hat^2 * d^2,     g_hat = prod_j c_j of the Givens sequence,
// because n_xi = Z t12 + z_new / g_hat (last coordinate) and Schur complements scale by the square of the
// diagonal entry of a triangular change of basis.  So the O(N^2) state (Phi, Q, Z) collapses to the inverse
// Cholesky factor of A (m x m, packed), two p x m coefficient blocks and R -- all resident in shared memory
// for the benchmark shapes -- and the per-candidate work drops from four N-GEMVs over L2/HBM to one
// triangular mat-vec in shared memory.  Instances that do not qualify (N0 != p, rank-deficient Pi_0) are
// marked n_r4 = -1 and handled by the literal kernel above.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int tri(int r) { return (r * (r + 1)) >> 1; }

// D (8 x 8) += A (8 x 4) B (4 x 8) on the FP64 tensor path: lane l holds A[l / 4][l % 4], B[l % 4][l / 4], D[l / 4][2 (l % 4) + {0, 1}]
__device__ __forceinline__ void dmma884_sel(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
// Row stride of the block's a- and t-vectors in shared memory: = 4 (mod 16) doubles, so that the 8 rows x 4 columns a tensor-path operand
// load touches fall into 2 x 16 distinct banks (two wavefronts, the minimum for 256 bytes)
__host__ __device__ __forceinline__ int round4_block_row_stride(int MM) { return ((MM + 11) / 16) * 16 + 4; }

template <int T, bool SMEM, int NT>
__global__ void __launch_bounds__(NT) round4_block_kernel(Round4Params P) {
    // Blocked variant of the shared-memory round 4: T candidates are evaluated together against the factorisation as it was
    // before the block (a rejected candidate never changes the state, so this speculation cannot fail); the few quantities
    // that depend on which block members were accepted -- the extra Cholesky entries e_ij, the pivots d_j^2 and the
    // leverages -- are resolved by an O(T^3) scalar "panel" step, exactly the recurrences of a blocked left-looking
    // Cholesky.  Barriers per candidate drop from ~4 to ~9/T and every phase has T times more parallel work.
    extern __shared__ double smem[];
    // NT = 256 when the state lives in shared memory; 1024 when it lives in global memory (large n or many model points):
    // every phase is then a stream of L2 accesses and four times the warps hide four times the latency.
    constexpr int nt = NT, nwarps = NT / 32;
    const int b = blockIdx.x, n = P.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NM = P.NM;
    const int deg = P.cfg.polynomial_degree;
    const int p = poly_dim(n, deg);
    const int pl = p > 0 ? p : 1, pb = pl | 1;
    const int MM = (NM - p) > 1 ? (NM - p) : 1;
    const int MS = round4_block_row_stride(MM);     // row stride of AV / TV
    // block buffers
    double* XI = smem;                 // T x n    candidate sites
    double* PH = XI + T * n;           // T x NM   kernel columns against the current centres
    double* CV = PH + T * NM;          // T x pl   c_xi
    double* UB = CV + T * pl;          // T x pl   u_xi = b_xi - Phi00 c_xi
    double* HV = UB + T * pl;          // T x pl   H pi_xi
    double* AV = HV + T * pl;          // T x MS   a_xi
    double* TV = AV + T * MS;          // T x MS   t_xi = L^{-1} a_xi
    double* Kx = TV + T * MS;          // T x T    phi(xi_i, xi_j)
    double* Ax = Kx + T * T;           // T x T    A_ij (i < j), A_jj on the diagonal
    double* Dx = Ax + T * T;           // T x T    t_i . t_j
    double* Sx = Dx + T * T;           // T x T    pi_i' H pi_j (i < j), leverage on the diagonal
    double* Ex = Sx + T * T;           // T x T    panel: e_ij (extra Cholesky entries), d_j on the diagonal
    double* Px = Ex + T * T;           // T x T    panel: inverse of the accepted part of [e, d]
    double* Sc = Px + T * T;           // T x T    panel: corrected s'_ij, 1 + lev'_j on the diagonal
    double* PT = Sc + T * T;           // T x pl   pi~ of the block members
    double* RI = PT + T * pl;          // T        1 / (1 + lev'_j) of accepted members
    double* tnp = RI + T;              // nwarps x T  per-warp partial ||t_j||^2 (room for 32 warps)
    double* tauq = tnp + 32 * T;       // pl
    double* red = tauq + pl;           // 80
    int* ib = reinterpret_cast<int*>(red + 80);   // ids[T], acc[T], pos[T], per-warp counts[nwarps], na
    double* st = red + 80 + 2 * T + 4 + 16;
    double* fs;
    if constexpr (SMEM) fs = st; else fs = (P.keep_fs ? P.keep_fs : P.fs) + (size_t)b * P.fs_stride;
    double* Ct = fs;                   // NM x n  coordinate-major centres
    double* M0 = Ct + NM * n;          // p x p   Pi_0^{-T}
    double* P00 = M0 + pl * pl;        // p x p   Phi(S0, S0)
    double* H = P00 + pl * pl;         // p x p   (Pi' Pi)^{-1} of the current point set
    double* Aq = H + pl * pl;          // p x p   scratch (QR of Pi_0)
    double* Qx = Aq + pl * pl;         // p x p   scratch (explicit Q_0)
    double* Tm = Qx + pl * pl;         // p x p   scratch (R_0^{-1})
    double* Gm = Tm + pl * pl;         // pb x MM  g_eta = Phi(S0, eta) - Phi00 c_eta
    double* Cm = Gm + pb * MM;         // pb x MM  c_eta
    double* Li = Cm + pb * MM;         // packed lower triangle of L^{-1}, row r at tri(r)

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
    if (!(N0 < max_points) || N0 > NM) { if (tid == 0) { P.n_r4[b] = 0; if (P.status) P.status[b] = (N0 > NM) ? -1 : 0; } return; }
    if (p > 0 && N0 != p) { if (tid == 0) P.n_r4[b] = -1; return; }         // literal kernel takes over

    // compact, ascending candidate list (results_in_box_indices minus the found set)
    unsigned char* cand = P.cand + (size_t)P.B * P.db_stride * sizeof(int) + (size_t)b * P.db_stride;
    int* clist = reinterpret_cast<int*>(P.cand) + (size_t)b * P.db_stride;
    for (int id = tid; id < n_db; id += nt) {
        bool ok = in_box(sites + (size_t)id * n, lb2, ub2, n);
        for (int f = 0; f < nf_ids && ok; ++f) ok = (found[f] != id + 1);
        cand[id] = ok ? 1 : 0;
    }
    for (int e = tid; e < N0 * n; e += nt) {
        int i = e / n, k = e % n;
        Ct[k * NM + i] = (i < nf_ids) ? sites[(size_t)(found[i] - 1) * n + k] : extra[(size_t)(i - nf_ids) * n + k];
    }
    if (tid == 0) red[76] = 0.0;
    __syncthreads();
    // The polynomial basis is centred at the first found point and scaled by the spread of S0: c_xi and the
    // leverage behind g_hat are invariant under that change of basis, and Pi_0 stays well conditioned for tiny Delta.
    double inv_s = 1.0;
    if (p > 1) {
        double mx = 0.0;
        for (int e = tid; e < N0 * n; e += nt) { int i = e / n, k = e % n; mx = fmax(mx, fabs(Ct[k * NM + i] - Ct[k * NM])); }
        mx = warp_max(mx);
        if (lane == 0) red[40 + warp] = mx;
        __syncthreads();
        mx = 0.0;
        for (int w = 0; w < nwarps; ++w) mx = fmax(mx, red[40 + w]);
        inv_s = mx > 0.0 ? 1.0 / mx : 1.0;
        __syncthreads();
    }
    if (p > 0) {
        for (int e = tid; e < p * p; e += nt) {
            int i = e % p, j = e / p;
            if (i >= j) {                          // symmetric (bit for bit: (a - b)^2 == (b - a)^2): lower half, mirrored
                double r2 = 0.0;
#pragma unroll 4
                for (int k = 0; k < n; ++k) { double d = Ct[k * NM + i] - Ct[k * NM + j]; r2 = fma(d, d, r2); }
                const double ph = rad_phi(P.rf, r2);
                P00[i + j * pl] = ph; P00[j + i * pl] = ph;
            }
            Aq[i + j * pl] = (j == 0) ? 1.0 : (Ct[(j - 1) * NM + i] - Ct[(j - 1) * NM]) * inv_s;
        }
        __syncthreads();
        // Pi_0^{-1} by Gauss-Jordan with partial pivoting on [Pi_0 | I] (p x 2p, Aq and Qx are adjacent), block parallel
        for (int e = tid; e < p * p; e += nt) { int i = e % p, j = e / p; Qx[i + j * pl] = (i == j) ? 1.0 : 0.0; }
        __syncthreads();
        for (int kk = 0; kk < p; ++kk) {
            // One barrier per step and no serial section: EVERY warp finds the pivot of column kk for itself (identical arithmetic,
            // hence identical result), and the warp that owns a column performs that column's row swap and scaling on the fly.
            // Column kk itself is dead after this step, so it is neither swapped nor scaled: its multipliers are read through
            // the swap (row piv <- the old diagonal entry).
            ArgMax mine; mine.v = 0.0; mine.id = -1;
            for (int i = kk + lane; i < p; i += 32) { ArgMax c_; c_.v = fabs(Aq[i + kk * pl]); c_.id = i; mine = better(mine, c_); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ArgMax t_; t_.v = __shfl_xor_sync(0xffffffffu, mine.v, o); t_.id = __shfl_xor_sync(0xffffffffu, mine.id, o);
                mine = better(mine, t_);
            }
            if (!(mine.v > 1e-12)) { if (tid == 0) P.n_r4[b] = -1; return; }       // Pi_0 (scaled to O(1)) is rank deficient: uniform exit
            const int piv = mine.id;
            const double rp = 1.0 / Aq[piv + kk * pl];
            const double dkk = Aq[kk + kk * pl];                // multiplier of row piv after the swap
            for (int c = kk + 1 + warp; c < 2 * p; c += nwarps) {
                const double a = Aq[piv + c * pl], bq = Aq[kk + c * pl];
                const double pk = a * rp;
                __syncwarp();
                if (pk == 0.0) {                                 // identity columns no pivot row has touched yet: only the swap
                    if (piv != kk && lane == 0) { Aq[piv + c * pl] = bq; Aq[kk + c * pl] = 0.0; }
                    continue;
                }
#pragma unroll 8
                for (int i = lane; i < p; i += 32) {
                    double v_;
                    if (i == kk) v_ = pk;
                    else if (i == piv) v_ = fma(-dkk, pk, bq);
                    else v_ = fma(-Aq[i + kk * pl], pk, Aq[i + c * pl]);
                    Aq[i + c * pl] = v_;
                }
            }
            __syncthreads();
        }
        for (int e = tid; e < p * p; e += nt) {    // M0 = Pi_0^{-T};  H = (Pi_0' Pi_0)^{-1} = Pi_0^{-1} Pi_0^{-T}
            const int r = e % p, c = e / p;
            M0[r + c * pl] = Qx[c + r * pl];
            if (r >= c) {                          // symmetric, and bit-identical under the swap of the two factors
                double a = 0.0;
#pragma unroll 4
                for (int k2 = 0; k2 < p; ++k2) a = fma(Qx[r + k2 * pl], Qx[c + k2 * pl], a);
                H[r + c * pl] = a; H[c + r * pl] = a;
            }
        }
    }
    const double phi0 = rad_phi(P.rf, 0.0);
    const double thr = P.chol_thr;
    const int base = (p > 0) ? p : N0;             // index of the first round-4 point among the centres
    int N = N0, m = 0, nr4 = 0;

    // ---- candidate list: warp w compacts the contiguous id segment [w L, (w+1) L)
    int nc;
    {
        const int L = (((n_db + nwarps - 1) / nwarps) + 31) & ~31;
        const int lo = warp * L, hi = min(n_db, lo + L);
        int cnt = 0;
        for (int i0 = lo; i0 < hi; i0 += 32) {
            const int id = i0 + lane;
            const unsigned msk = __ballot_sync(0xffffffffu, id < hi && cand[id]);
            cnt += __popc(msk);
        }
        if (lane == 0) ib[3 * T + warp] = cnt;
        __syncthreads();
        int basew = 0;
        for (int w = 0; w < warp; ++w) basew += ib[3 * T + w];
        nc = 0;
        for (int w = 0; w < nwarps; ++w) nc += ib[3 * T + w];
        for (int i0 = lo; i0 < hi; i0 += 32) {
            const int id = i0 + lane;
            const bool f = id < hi && cand[id];
            const unsigned msk = __ballot_sync(0xffffffffu, f);
            if (f) clist[basew + __popc(msk & ((1u << lane) - 1u))] = id;
            basew += __popc(msk);
        }
        __syncthreads();
    }

    for (int pos = 0; pos < nc && N < max_points && nr4 < P.r4_stride; pos += T) {
        const int tb = min(T, nc - pos);
        // ---- P0: candidate sites
        for (int e = tid; e < tb * n; e += nt) { const int j = e / n, k = e % n; XI[j * n + k] = sites[(size_t)clist[pos + j] * n + k]; }
        if (tid < tb) ib[tid] = clist[pos + tid];
        __syncthreads();
        for (int e = tid; e < tb * p; e += nt) { const int j = e / p, c = e % p; PT[j * pl + c] = (c == 0) ? 1.0 : (XI[j * n + c - 1] - Ct[(c - 1) * NM]) * inv_s; }
        __syncthreads();
        // ---- P1: leverage vectors, Lagrange coefficients, kernel columns: one thread per row, all block members at once
        // (the matrix row / centre is loaded once and feeds T independent accumulator chains)
        for (int r = tid; r < 2 * p + N; r += nt) {
            double acc[T];
#pragma unroll
            for (int j = 0; j < T; ++j) acc[j] = 0.0;
            if (r < 2 * p) {
                const double* Mx = (r < p) ? (H + r) : (M0 + (r - p));
#pragma unroll 8
                for (int c = 0; c < p; ++c) {
                    const double mv = Mx[c * pl];
#pragma unroll
                    for (int j = 0; j < T; ++j) acc[j] = fma(mv, PT[j * pl + c], acc[j]);
                }
                double* dst = (r < p) ? (HV + r) : (CV + (r - p));
#pragma unroll
                for (int j = 0; j < T; ++j) if (j < tb) dst[j * pl] = acc[j];
            } else {
                const int i = r - 2 * p;
#pragma unroll 8
                for (int k = 0; k < n; ++k) {
                    const double cv_ = Ct[k * NM + i];
#pragma unroll
                    for (int j = 0; j < T; ++j) { const double d = XI[j * n + k] - cv_; acc[j] = fma(d, d, acc[j]); }
                }
#pragma unroll
                for (int j = 0; j < T; ++j) if (j < tb) PH[j * NM + i] = rad_phi(P.rf, acc[j]);
            }
        }
        for (int q = tid; q < tb * tb; q += nt) {
            const int i = q / tb, j2 = q % tb;
            if (i < j2) {
                double r2 = 0.0;
                for (int k = 0; k < n; ++k) { double d = XI[i * n + k] - XI[j2 * n + k]; r2 = fma(d, d, r2); }
                Kx[i * T + j2] = rad_phi(P.rf, r2);
            }
        }
        __syncthreads();
        // ---- P2: u_j = b_j - Phi00 c_j ; a_j[eta] = phi(eta, xi_j) - g_eta.c_j - c_eta.b_j ; leverage l_j = pi_j.h_j
        if constexpr (!SMEM) {
            // a_j[eta] for 8 values of eta per warp on the FP64 tensor path: K runs over the p entries of g_eta (against c_j) and the p
            // entries of c_eta (against b_j = the first p kernel values of member j)
            constexpr int NJT = (T + 7) / 8;
            const int lr = lane >> 2, lk = lane & 3;
            const int ntile = (m + 7) >> 3;
            for (int t = warp; t < ntile; t += nwarps) {
                const int eta = (t << 3) + lr;
                const bool eok = eta < m;
                const double* ge = Gm + (eok ? eta : 0) * pb; const double* ce = Cm + (eok ? eta : 0) * pb;
                double d[NJT][2];
#pragma unroll
                for (int jt = 0; jt < NJT; ++jt) d[jt][0] = d[jt][1] = 0.0;
#pragma unroll 2
                for (int c0 = 0; c0 < p; c0 += 4) {
                    const int c = c0 + lk;
                    const bool cok = c < p;
                    const double ag = (eok && cok) ? ge[c] : 0.0, ac = (eok && cok) ? ce[c] : 0.0;
#pragma unroll
                    for (int jt = 0; jt < NJT; ++jt) {
                        const bool jok = cok && (jt * 8 + lr < T);
                        dmma884_sel(d[jt], ag, jok ? CV[(jt * 8 + lr) * pl + c] : 0.0);
                        dmma884_sel(d[jt], ac, jok ? PH[(jt * 8 + lr) * NM + c] : 0.0);
                    }
                }
#pragma unroll
                for (int jt = 0; jt < NJT; ++jt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = jt * 8 + 2 * lk + e;
                        if (j < tb && eok) AV[j * MS + eta] = PH[j * NM + base + eta] - d[jt][e];
                    }
            }
        }
        for (int r = tid; r < (SMEM ? p + m : p); r += nt) {
            double acc[T];
#pragma unroll
            for (int j = 0; j < T; ++j) acc[j] = 0.0;
            if (r < p) {
#pragma unroll 8
                for (int c = 0; c < p; ++c) {
                    const double mv = P00[r + c * pl];
#pragma unroll
                    for (int j = 0; j < T; ++j) acc[j] = fma(mv, CV[j * pl + c], acc[j]);
                }
#pragma unroll
                for (int j = 0; j < T; ++j) if (j < tb) UB[j * pl + r] = PH[j * NM + r] - acc[j];
            } else {
                const int eta = r - p;
                const double* ge = Gm + eta * pb; const double* ce = Cm + eta * pb;
#pragma unroll 8
                for (int c = 0; c < p; ++c) {
                    const double gv = ge[c], cv_ = ce[c];
#pragma unroll
                    for (int j = 0; j < T; ++j) acc[j] = fma(gv, CV[j * pl + c], fma(cv_, PH[j * NM + c], acc[j]));
                }
#pragma unroll
                for (int j = 0; j < T; ++j) if (j < tb) AV[j * MS + eta] = PH[j * NM + base + eta] - acc[j];
            }
        }
        if (tid >= nt - 32 && lane < tb) {         // leverages by the last warp
            const double* hv = HV + lane * pl; const double* pt = PT + lane * pl;
            double a = 0.0;
            for (int c = 0; c < p; ++c) a = fma(hv[c], pt[c], a);
            Sx[lane * T + lane] = a;
        }
        __syncthreads();
        // ---- P3: t_j = L^{-1} a_j for all block members at once (G threads per row share the row of L^{-1})
        const int G = (!SMEM) ? 8 : ((m > 64) ? 2 : ((m > 32) ? 4 : 8));   // state in global memory: eight lanes read a 64-byte piece of a row of L^{-1}
        const int rows_per_pass = nt / G;
        if constexpr (!SMEM) {
            // State in global memory (large systems): T = L^{-1} A' as FP64 tensor-path tiles.  A warp owns 8 rows of L^{-1} at a time
            // (mma.m8n8k4: the 8 x 4 piece of L^{-1} straight from L2 / HBM as the A operand, 4 x 8 pieces of the block's a-vectors from
            // shared memory as B) -- one shared-memory operand per 256 FMA where the lane-per-column loop below reads one per FMA, which
            // is what bounded this phase (ncu: 29 % of the kernel's shared-memory wavefronts).  Row tiles are dealt to the warps in a
            // snake so that the triangular row lengths balance.
            constexpr int NJT = (T + 7) / 8;
            const int lr = lane >> 2, lk = lane & 3;
            const int ntile = (m + 7) >> 3;
            double tn2[NJT][2];
#pragma unroll
            for (int jt = 0; jt < NJT; ++jt) tn2[jt][0] = tn2[jt][1] = 0.0;
            for (int i = 0; i * nwarps < ntile; ++i) {
                const int t = i * nwarps + ((i & 1) ? (nwarps - 1 - warp) : warp);
                if (t >= ntile) continue;
                const int r = (t << 3) + lr;
                const bool rok = r < m;
                const double* lrow = Li + tri(rok ? r : 0);
                const int cend = min((t << 3) + 8, m);
                double d[NJT][2];
#pragma unroll
                for (int jt = 0; jt < NJT; ++jt) d[jt][0] = d[jt][1] = 0.0;
#pragma unroll 8
                for (int c0 = 0; c0 < cend; c0 += 4) {
                    const int c = c0 + lk;
                    const double a = (rok && c <= r) ? lrow[c] : 0.0;
#pragma unroll
                    for (int jt = 0; jt < NJT; ++jt) {
                        const double bv = (c < m && jt * 8 + lr < T) ? AV[(jt * 8 + lr) * MS + c] : 0.0;
                        dmma884_sel(d[jt], a, bv);
                    }
                }
#pragma unroll
                for (int jt = 0; jt < NJT; ++jt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int j = jt * 8 + 2 * lk + e;
                        if (j < T && rok) TV[j * MS + r] = d[jt][e];
                        tn2[jt][e] = fma(d[jt][e], d[jt][e], tn2[jt][e]);          // rows beyond m carry zeros
                    }
            }
#pragma unroll
            for (int jt = 0; jt < NJT; ++jt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    double v = tn2[jt][e];
                    v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
                    const int j = jt * 8 + 2 * lk + e;
                    if (lr == 0 && j < T) tnp[warp * T + j] = v;
                }
        } else {
            double tn[T];
#pragma unroll
            for (int j = 0; j < T; ++j) tn[j] = 0.0;
            for (int r0 = 0; r0 < m; r0 += rows_per_pass) {
                const int r = r0 + tid / G, l = tid % G;
                double acc[T];
#pragma unroll
                for (int j = 0; j < T; ++j) acc[j] = 0.0;
                if (r < m) {
                    const double* lr = Li + tri(r);
#pragma unroll 4
                    for (int c = l; c <= r; c += G) {
                        const double lv = lr[c];
#pragma unroll
                        for (int j = 0; j < T; ++j) acc[j] = fma(lv, AV[j * MS + c], acc[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < T; ++j) {
                    double a = acc[j];
                    for (int o = G >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
                    if (r < m && l == 0) { TV[j * MS + r] = a; tn[j] = fma(a, a, tn[j]); }
                }
            }
#pragma unroll
            for (int j = 0; j < T; ++j) { const double s_ = warp_sum(tn[j]); if (lane == 0) tnp[warp * T + j] = s_; }
        }
        __syncthreads();
        // ---- P4: pair quantities, 8 lanes per pair (i <= j): A_ij, t_i.t_j, pi_i' H pi_j ; diagonal A_jj
        {
            const int npair = (tb * (tb + 1)) / 2;
            for (int q0 = 0; q0 < npair; q0 += nt / 8) {
                const int q = q0 + (tid >> 3), l8 = tid & 7;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0;
                int i = 0, j = 0;
                if (q < npair) {
                    while (((j + 1) * (j + 2)) / 2 <= q) ++j;          // q = j (j + 1) / 2 + i,  i <= j
                    i = q - (j * (j + 1)) / 2;
                    if (i < j) {
                        for (int c = l8; c < p; c += 8) {
                            a0 = fma(UB[i * pl + c], CV[j * pl + c], a0);
                            a0 = fma(CV[i * pl + c], PH[j * NM + c], a0);
                            a2 = fma(HV[i * pl + c], PT[j * pl + c], a2);
                        }
                        for (int r = l8; r < m; r += 8) a1 = fma(TV[i * MS + r], TV[j * MS + r], a1);
                    } else {
                        for (int c = l8; c < p; c += 8) a0 = fma(CV[j * pl + c], PH[j * NM + c] + UB[j * pl + c], a0);
                    }
                }
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) {
                    a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); a2 += __shfl_xor_sync(0xffffffffu, a2, o);
                }
                if (q < npair && l8 == 0) {
                    if (i < j) { Ax[i * T + j] = Kx[i * T + j] - a0; Dx[i * T + j] = a1; Sx[i * T + j] = a2; }
                    else Ax[j * T + j] = phi0 - a0;
                }
            }
        }
        __syncthreads();
        // ---- P5: panel (one warp, lane = block member): resolve the dependence on accepted block members.
        // Lane i decides once every accepted l < i has been folded into its pivot and leverage; the lanes j > i then
        // fold member i into theirs (the recurrences of a left-looking Cholesky / of successive Sherman-Morrison updates).
        // (state in global memory) While warp 0 walks the panel, the other warps already form the products  t_j' L^{-1}  of ALL block
        // members -- they do not depend on which members get accepted -- and park them in the rows m + j of L^{-1}, which are free
        // until the block is appended; after the barrier a thread per column applies the accepted part of P_^{-1} in place.  Needs the
        // rows m .. m + tb - 1 to exist, i.e. not within tb of the cap; otherwise the products are formed after the panel (P6 below).
        bool early = false;
        if constexpr (!SMEM) early = (m > 0) && (m + tb <= MM);
        if (warp != 0 && early) {
            if constexpr (!SMEM) {
                constexpr int NJT = (T + 7) / 8;
                const int lr = lane >> 2, lk = lane & 3, nw = nwarps - 1, wi = warp - 1;
                const int ntile = (m + 7) >> 3;
                for (int i = 0; i * nw < ntile; ++i) {
                    const int t = i * nw + ((i & 1) ? (nw - 1 - wi) : wi);
                    if (t >= ntile) continue;
                    const int c0 = t << 3, c = c0 + lr;
                    double d[NJT][2];
#pragma unroll
                    for (int jt = 0; jt < NJT; ++jt) d[jt][0] = d[jt][1] = 0.0;
#pragma unroll 8
                    for (int r0 = c0; r0 < m; r0 += 4) {
                        const int r = r0 + lk;
                        const bool rok = r < m;
                        const double bv = (rok && c <= r) ? Li[tri(r) + c] : 0.0;
#pragma unroll
                        for (int jt = 0; jt < NJT; ++jt) {
                            const double a = (rok && jt * 8 + lr < tb) ? TV[(jt * 8 + lr) * MS + r] : 0.0;
                            dmma884_sel(d[jt], a, bv);
                        }
                    }
#pragma unroll
                    for (int jt = 0; jt < NJT; ++jt)
                        if (jt * 8 + lr < tb) {
#pragma unroll
                            for (int e = 0; e < 2; ++e) { const int cc = c0 + 2 * lk + e; if (cc < m) Li[tri(m + jt * 8 + lr) + cc] = d[jt][e]; }
                        }
                }
            }
        }
        if (warp == 0) {
            const int j = lane;
            double dj2 = 0.0, lev = 0.0;
            if (j < tb) {
                double tnj = 0.0;
                for (int w = 0; w < nwarps; ++w) tnj += tnp[w * T + j];
                dj2 = Ax[j * T + j] - tnj; lev = Sx[j * T + j];
            }
            int na_ = 0;
            int al[T];
            for (int i = 0; i < tb; ++i) {
                const double d2i = __shfl_sync(0xffffffffu, dj2, i), levi = __shfl_sync(0xffffffffu, lev, i);
                const double tau2 = d2i / (1.0 + levi);          // == sigma - ||L^-1 v||^2 of RbfModel.jl:447-449
                const bool ok = (tau2 > thr) && (N + na_ < max_points) && (nr4 + na_ < P.r4_stride);    // RbfModel.jl:452, 402
                if (lane == 0) ib[T + i] = ok ? 1 : 0;
                if (ok) {
                    const double di = sqrt(d2i), ri = 1.0 / (1.0 + levi), rdi = 1.0 / di;
                    if (lane == 0) { ib[2 * T + na_] = i; Ex[i * T + i] = di; Sc[i * T + i] = 1.0 + levi; RI[i] = ri; }
                    if (j > i && j < tb) {
                        double e = Ax[i * T + j] - Dx[i * T + j];
                        double sp = Sx[i * T + j];
#pragma unroll
                        for (int ql = 0; ql < T; ++ql) if (ql < na_) {
                            const int l = al[ql];
                            e = fma(-Ex[l * T + i], Ex[l * T + j], e);
                            sp = fma(-Sc[l * T + i] * RI[l], Sc[l * T + j], sp);
                        }
                        e *= rdi;
                        Ex[i * T + j] = e; Sc[i * T + j] = sp;
                        dj2 = fma(-e, e, dj2);
                        lev = fma(-sp * ri, sp, lev);
                    }
#pragma unroll
                    for (int ql = 0; ql < T; ++ql) if (ql == na_) al[ql] = i;
                    ++na_;
                }
                __syncwarp();
            }
            // inverse of the accepted panel P_ = [e_ij below the diagonal, d_j on it]: lane c owns column c
            if (lane < na_) {
                const int c = lane;
                for (int q = c; q < na_; ++q) {
                    const int jq = ib[2 * T + q];
                    double v = (c == q) ? 1.0 : 0.0;
                    for (int q2 = c; q2 < q; ++q2) v = fma(-Ex[ib[2 * T + q2] * T + jq], Px[q2 * T + c], v);
                    Px[q * T + c] = v / Ex[jq * T + jq];
                }
            }
            if (lane == 0) ib[3 * T + nwarps] = na_;
        }
        __syncthreads();
        const int na = ib[3 * T + nwarps];
        // ---- P6: append the accepted members
        if (na > 0) {
            // new rows of L^{-1}: columns < m from  -P_^{-1} (T_ L^{-1}),  columns >= m from P_^{-1}
            if (early) {
                for (int c = tid; c < m; c += nt) {
                    double xs[T];
#pragma unroll
                    for (int q2 = 0; q2 < T; ++q2) xs[q2] = (q2 < na) ? Li[tri(m + ib[2 * T + q2]) + c] : 0.0;     // all reads of the column first
#pragma unroll
                    for (int q = 0; q < T; ++q) if (q < na) {
                        double v = 0.0;
#pragma unroll
                        for (int q2 = 0; q2 < T; ++q2) if (q2 <= q) v = fma(Px[q * T + q2], xs[q2], v);
                        Li[tri(m + q) + c] = -v;
                    }
                }
            } else if constexpr (!SMEM) {
                // tensor-path twin of the loop below: a warp owns 8 columns of L^{-1}; A = the accepted members' t-vectors (shared memory),
                // B = 4 x 8 pieces of L^{-1} below the diagonal; the T x T triangular factor P_^{-1} is applied from registers through
                // shuffles (lane (q, cpair) collects column pair cpair of every row q2 <= q)
                constexpr int NJT = (T + 7) / 8;
                const int lr = lane >> 2, lk = lane & 3;
                const int ntile = (m + 7) >> 3;
                int jq[NJT];
#pragma unroll
                for (int jt = 0; jt < NJT; ++jt) jq[jt] = (jt * 8 + lr < na) ? ib[2 * T + jt * 8 + lr] : -1;
                for (int i = 0; i * nwarps < ntile; ++i) {
                    const int t = i * nwarps + ((i & 1) ? (nwarps - 1 - warp) : warp);
                    if (t >= ntile) continue;
                    const int c0 = t << 3, c = c0 + lr;
                    double d[NJT][2];
#pragma unroll
                    for (int jt = 0; jt < NJT; ++jt) d[jt][0] = d[jt][1] = 0.0;
#pragma unroll 4
                    for (int r0 = c0; r0 < m; r0 += 4) {
                        const int r = r0 + lk;
                        const bool rok = r < m;
                        const double bv = (rok && c <= r) ? Li[tri(r) + c] : 0.0;        // (c <= r < m implies c < m)
#pragma unroll
                        for (int jt = 0; jt < NJT; ++jt) {
                            const double a = (rok && jq[jt] >= 0) ? TV[jq[jt] * MS + r] : 0.0;
                            dmma884_sel(d[jt], a, bv);
                        }
                    }
                    double v[NJT][2];
#pragma unroll
                    for (int jt = 0; jt < NJT; ++jt) v[jt][0] = v[jt][1] = 0.0;
#pragma unroll
                    for (int q2 = 0; q2 < T; ++q2) {
                        const int src = ((q2 & 7) << 2) | lk;
                        const double x0 = __shfl_sync(0xffffffffu, d[q2 >> 3][0], src), x1 = __shfl_sync(0xffffffffu, d[q2 >> 3][1], src);
#pragma unroll
                        for (int jt = 0; jt < NJT; ++jt) {
                            const int q = jt * 8 + lr;
                            if (q2 <= q && q < na) { const double pv = Px[q * T + q2]; v[jt][0] = fma(pv, x0, v[jt][0]); v[jt][1] = fma(pv, x1, v[jt][1]); }
                        }
                    }
#pragma unroll
                    for (int jt = 0; jt < NJT; ++jt) {
                        const int q = jt * 8 + lr;
                        if (q < na) {
#pragma unroll
                            for (int e = 0; e < 2; ++e) { const int cc = c0 + 2 * lk + e; if (cc < m) Li[tri(m + q) + cc] = -v[jt][e]; }
                        }
                    }
                }
            } else
            for (int c0 = 0; c0 < m; c0 += rows_per_pass) {
                const int c = c0 + tid / G, l = tid % G;
                double acc[T];
#pragma unroll
                for (int q = 0; q < T; ++q) acc[q] = 0.0;
                if (c < m)
#pragma unroll 4
                    for (int r = c + l; r < m; r += G) {
                        const double lv = Li[tri(r) + c];
#pragma unroll
                        for (int q = 0; q < T; ++q) if (q < na) acc[q] = fma(TV[ib[2 * T + q] * MS + r], lv, acc[q]);
                    }
#pragma unroll
                for (int q = 0; q < T; ++q) for (int o = G >> 1; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
                if (c < m && l == 0) {
#pragma unroll
                    for (int q = 0; q < T; ++q) if (q < na) {
                        double v = 0.0;
#pragma unroll
                        for (int q2 = 0; q2 < T; ++q2) if (q2 <= q) v = fma(Px[q * T + q2], acc[q2], v);
                        Li[tri(m + q) + c] = -v;
                    }
                }
            }
            for (int e = tid; e < na * na; e += nt) { const int q = e / na, c = e % na; if (c <= q) Li[tri(m + q) + m + c] = Px[q * T + c]; }
            for (int e = tid; e < na * p; e += nt) { const int q = e / p, r = e % p, j = ib[2 * T + q]; Gm[(m + q) * pb + r] = UB[j * pl + r]; Cm[(m + q) * pb + r] = CV[j * pl + r]; }
            for (int e = tid; e < na * n; e += nt) { const int q = e / n, k = e % n; Ct[k * NM + N + q] = XI[ib[2 * T + q] * n + k]; }
            if (tid < na) r4[nr4 + tid] = ib[ib[2 * T + tid]] + 1;
            // corrected leverage vectors h'_q = h_q - sum_{q' < q} h'_q' s'_q'q / (1 + lev'_q'), then H -= sum h' h'^T / (1 + lev')
            for (int r = tid; r < p; r += nt)
                for (int q = 1; q < na; ++q) {
                    const int jq = ib[2 * T + q];
                    double h = HV[jq * pl + r];
                    for (int q2 = 0; q2 < q; ++q2) { const int j2 = ib[2 * T + q2]; h = fma(-HV[j2 * pl + r], Sc[j2 * T + jq] * RI[j2], h); }
                    HV[jq * pl + r] = h;
                }
            __syncthreads();
#pragma unroll 4
            for (int e = tid; e < p * p; e += nt) {
                const int a_ = e % p, b_ = e / p;
                double h = H[a_ + b_ * pl];
                for (int q = 0; q < na; ++q) { const int jq = ib[2 * T + q]; h = fma(-HV[jq * pl + a_] * RI[jq], HV[jq * pl + b_], h); }
                H[a_ + b_ * pl] = h;
            }
            N += na; m += na; nr4 += na;
        }
        __syncthreads();
    }
    if (tid == 0) { P.n_r4[b] = nr4; if (P.status) P.status[b] = 0; }
    if (P.keep_fs) {
        // keep the factorisation for mrbf_build_prepared: centres, Pi_0^{-T}, g/c blocks, packed L^{-1}
        double* out = P.keep_fs + (size_t)b * P.fs_stride;
        if constexpr (SMEM) {
            const int used_c = pb * m, used_l = tri(m);
            for (int e = tid; e < NM * n; e += nt) { int i = e % NM; if (i < N) out[e] = Ct[e]; }
            double* o = out + NM * n;
            for (int e = tid; e < pl * pl; e += nt) o[e] = M0[e];
            o = out + (Gm - fs);
            for (int e = tid; e < used_c; e += nt) { o[e] = Gm[e]; o[pb * MM + e] = Cm[e]; }
            o = out + (Li - fs);
            for (int e = tid; e < used_l; e += nt) o[e] = Li[e];
        }
        if (tid == 0) { out[P.fs_stride - 1] = inv_s; out[P.fs_stride - 2] = (double)N0; out[P.fs_stride - 3] = (double)m; P.elig[b] = 1; }
    }
}

size_t round4_fast_state_doubles(int n, int NM, int p) {
    int pl = p > 0 ? p : 1, pb = pl | 1; int MM = (NM - p) > 1 ? (NM - p) : 1;
    return (size_t)NM * n + 6 * (size_t)pl * pl + 2 * (size_t)pb * MM + (size_t)MM * (MM + 1) / 2 + 8;
}
// offsets (in doubles) of the blocks mrbf_build_prepared reads from a kept state
void round4_fast_state_layout(int n, int NM, int p, size_t* off_M0, size_t* off_G, size_t* off_C, size_t* off_L) {
    int pl = p > 0 ? p : 1, pb = pl | 1; int MM = (NM - p) > 1 ? (NM - p) : 1;
    *off_M0 = (size_t)NM * n;
    *off_G = *off_M0 + 6 * (size_t)pl * pl;
    *off_C = *off_G + (size_t)pb * MM;
    *off_L = *off_C + (size_t)pb * MM;
}

// ------------------------------------------------------------------------------------------------
// Training-set gather (_collect_indices order: centre, r1, r2, r3, r4), RbfModel.jl:178-186, 754-757
// ------------------------------------------------------------------------------------------------
__global__ void gather_training_kernel(GatherParams P) {
    const int b = blockIdx.x, n = P.n, k = P.k, tid = threadIdx.x, nt = blockDim.x;
    if (P.skip && P.skip[b]) return;
    const int n1 = P.n_r1[b], n2 = P.n_r2[b], n3 = P.n_r3[b], n4 = P.n_r4[b];
    const int N = 1 + n1 + n2 + n3 + n4;
    const double* sites = P.sites + (size_t)b * P.db_stride * n;
    const double* values = P.values + (size_t)b * P.db_stride * k;
    double* ts = P.train_sites + (size_t)b * P.train_stride * n;
    double* tv = P.train_values + (size_t)b * P.train_stride * k;
    if (tid == 0) P.N[b] = N <= P.train_stride ? N : -N;
    if (N > P.train_stride) return;
    for (int e = tid; e < N * (n + k); e += nt) {
        int i = e / (n + k), c = e % (n + k);
        int id = -1, j3 = -1;
        if (i == 0) id = P.x_index[b];
        else if (i < 1 + n1) id = P.r1[(size_t)b * n + i - 1];
        else if (i < 1 + n1 + n2) id = P.r2[(size_t)b * n + i - 1 - n1];
        else if (i < 1 + n1 + n2 + n3) j3 = i - 1 - n1 - n2;
        else id = P.r4[(size_t)b * P.r4_stride + i - 1 - n1 - n2 - n3];
        if (c < n) ts[(size_t)i * n + c] = (j3 >= 0) ? P.r3_sites[((size_t)b * n + j3) * n + c] : sites[(size_t)(id - 1) * n + c];
        else tv[(size_t)i * k + c - n] = (j3 >= 0) ? P.r3_values[((size_t)b * n + j3) * k + c - n] : values[(size_t)(id - 1) * k + c - n];
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
size_t select_smem_bytes(int n, bool wz_in_smem, int st_doubles, int db_stride) {
    size_t d = 8 * (size_t)n + 80 + 24 + 2 * (((size_t)db_stride + 3) / 4);
    if (wz_in_smem) d += 2 * (size_t)n * ((n + 3) & ~3) + (size_t)st_doubles;
    return d * sizeof(double);
}
size_t round4_vec_doubles(int n, int NM, int p) { int pl = p > 0 ? p : 1; return (size_t)n + 5 * (size_t)NM + 4 * (size_t)pl + 80; }
size_t round4_ws_doubles(int n, int NM, int p) {
    int pl = p > 0 ? p : 1;
    return (size_t)NM * n + (size_t)NM * NM + (size_t)NM * pl + (size_t)pl * pl + 2 * (size_t)NM * NM;
}

template <bool WZS, bool STS, int NT>
static cudaError_t launch_select_t(const SelectParams& P, size_t smem, cudaStream_t s) {
    cudaError_t e = raise_dyn_smem(select_rounds123_kernel<WZS, STS, NT>, smem);
    if (e != cudaSuccess) return e;
    select_rounds123_kernel<WZS, STS, NT><<<P.B, NT, smem, s>>>(P);
    return cudaGetLastError();
}
cudaError_t launch_select_rounds123(const SelectParams& P, size_t smem, cudaStream_t s) {
    if (P.wz_in_smem && P.st_in_smem) return launch_select_t<true, true, 256>(P, smem, s);
    if (P.wz_in_smem) return launch_select_t<true, false, 256>(P, smem, s);
    // W / Z in the global workspace (n > ~100): every phase streams L2, twice the warps hide twice the latency (the block arg-max
    // has scratch for 16 warps)
    return launch_select_t<false, false, 512>(P, smem, s);
}
cudaError_t launch_round4(const Round4Params& P, size_t smem, cudaStream_t s, int grid) {
    cudaError_t e = raise_dyn_smem(round4_kernel, smem);
    if (e != cudaSuccess) return e;
    round4_kernel<<<grid, 256, smem, s>>>(P);
    return cudaGetLastError();
}
size_t round4_block_vec_doubles(int T, int n, int NM, int p) {
    int pl = p > 0 ? p : 1; int MM = (NM - p) > 1 ? (NM - p) : 1;
    return (size_t)T * n + (size_t)T * NM + 4 * (size_t)T * pl + 2 * (size_t)T * round4_block_row_stride(MM) + 7 * (size_t)T * T + 33 * (size_t)T + pl + 80 + 2 * T + 4 + 16;
}
template <int T>
static cudaError_t launch_block_t(const Round4Params& P, size_t smem, cudaStream_t s) {
    cudaError_t e;
    if (P.fs_in_smem) {
        e = raise_dyn_smem(round4_block_kernel<T, true, 256>, smem);
        if (e != cudaSuccess) return e;
        round4_block_kernel<T, true, 256><<<P.B, 256, smem, s>>>(P);
    } else {
        e = raise_dyn_smem(round4_block_kernel<T, false, 1024>, smem);
        if (e != cudaSuccess) return e;
        round4_block_kernel<T, false, 1024><<<P.B, 1024, smem, s>>>(P);
    }
    return cudaGetLastError();
}
cudaError_t launch_round4_block(const Round4Params& P, int T, size_t smem, cudaStream_t s) {
    if (T == 16) {      // global-memory state, many accepted points: 16 candidates per pass over the packed L^{-1} (half the L2 / HBM traffic)
        cudaError_t e = raise_dyn_smem(round4_block_kernel<16, false, 512>, smem);
        if (e != cudaSuccess) return e;
        round4_block_kernel<16, false, 512><<<P.B, 512, smem, s>>>(P);
        return cudaGetLastError();
    }
    return T == 8 ? launch_block_t<8>(P, smem, s) : launch_block_t<4>(P, smem, s);
}
cudaError_t launch_gather_training(const GatherParams& P, cudaStream_t s) {
    gather_training_kernel<<<P.B, 128, 0, s>>>(P);
    return cudaGetLastError();
}

}  // namespace mrbf
