// Rounds 1-3 for the headline shape (n <= 32 variables, databases of <= 128 sites): the affinely-independent filter with its
// state in registers and the scoring GEMM on the FP64 tensor path (mma.sync m8n8k4 f64).
//
// Reference behaviour restated (never copied): src/models/AffinelyIndependentPoints.jl:4-106 (filter), src/models/RbfModel.jl:205-307,
// 518-655 (rounds 1-3 and their gating), src/Databases.jl:324-327 (box scan), src/utilities.jl:126-221 (wall step).
//
// Same mathematics as select_rounds123_kernel (mrbf_select.cu) -- trailing block W of the Householder Q updated by one LAPACK-style
// reflector per accepted point, projection coefficients y = W's of every candidate updated by the same reflector, scores
// || W D^-2 y ||_inf (D = column inf-norms of W) -- but laid out for the machine instead of for the formula:
//   * the reflector acts IN PLACE: W <- W H, y <- H y on the live columns j..n-1 and column j simply dies (no shift by one column
//     per step), so every matrix element keeps its owner thread for the whole run;
//   * y (32 x 128) lives in registers as the A fragments of the scoring product: warp w owns candidates 16w..16w+15, a thread
//     holds y[4 ks + lane%4][16 w + 8 ct + lane/4] for ks < 8, ct < 2 -- 16 doubles; the reflector update of y is 16 FMAs and two
//     quad shuffles per thread, no shared memory;
//   * W (32 x 32) lives in registers too, four consecutive columns of one row per thread; its update, the column maxima and the
//     scaled copy W D^-2 (the B operand, written to shared memory packed per k-step so that a fragment load is one contiguous
//     256-byte wavefront pair) cost ~100 instructions per thread and step;
//   * scores^T = y^T (W D^-2)^T: 8 DMMAs per warp and live k-step; the accumulator layout puts eight rows of ONE candidate in a
//     thread, so || . ||_inf is seven in-thread max operations and two quad shuffles;
//   * the warp winners publish their score together with their column of y, so the next reflector starts right after the one
//     barrier of the arg-max.  Three barriers per accepted point.
//   * the shifted seeds are never written to memory (the coordinate-major seed workspace of the general kernel, 126 MB per 4096
//     instances, is gone): y is loaded straight from the database rows when W is still the identity.
#include "mrbf_common.cuh"
#include "mrbf_kernels.h"

namespace mrbf {

#define CF_BOX1 1
#define CF_BOX2 2
#define CF_USED 4

namespace {

constexpr int NWARP = 8, NTHR = NWARP * 32, NCMAX = 16 * NWARP;     // 128 candidates per filter run

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

constexpr int WTS = 132;         // doubles per k-step of the packed B operand (128 + 4: the column owners' stores spread over the banks)

struct Smem {
    double x[32], lb1[32], ub1[32], lb2[32], ub2[32];
    double Wt[8 * WTS];          // W D^-2 packed [ks][row][4]; row-major scratch copy of W at run boundaries
    double cm[NWARP * 32];       // per-warp column maxima of |W|
    double xw[NWARP * 32];       // per-warp winner: the reflector vector v of its column of y
    double tauw[NWARP];          //                  tau of that reflector
    double wv[NWARP];            //                  score
    int wi[NWARP];               //                  position (-1: none)
    int cl[NCMAX];               // compacted candidate ids of the current run
    int ctl[8];
    unsigned char fl[NCMAX];     // flag bytes (box 1, box 2, picked)
};

struct Filter {
    double Y[8][2];              // projection coefficients of this thread's two candidates (A fragments)
    double w[4];                 // W[r][4g .. 4g+3]
    int jY;                      // accepted points so far = first live column
};

// max of two non-negative, NaN-free doubles.  fmax() costs seven instructions (NaN handling); IEEE doubles >= 0 order like their bit
// patterns, so this is two integer compares and two selects on the integer pipe -- the FP64 pipe is the contended one here.
#ifndef MRBF_MAXNN_FP
__device__ __forceinline__ double maxnn(double a, double b) {
    const unsigned long long ua = (unsigned long long)__double_as_longlong(a), ub = (unsigned long long)__double_as_longlong(b);
    return __longlong_as_double((long long)(ua > ub ? ua : ub));
}
#else
__device__ __forceinline__ double maxnn(double a, double b) { return a > b ? a : b; }
#endif
// max over the four lanes of a quad
__device__ __forceinline__ double quad_max(double v) {
    v = maxnn(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return maxnn(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ double quad_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// Warp arg-max of non-negative scores (first maximiser wins ties); result in every lane.
__device__ __forceinline__ void warp_argmax_pos(double v, int id, double& bv, int& bid) {
    if (id < 0) v = 0.0;
    const unsigned vh = (unsigned)__double2hiint(v), vl = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_max_sync(0xffffffffu, vh);
    const unsigned ml = __reduce_max_sync(0xffffffffu, vh == mh ? vl : 0u);
    const unsigned mid = __reduce_min_sync(0xffffffffu, (vh == mh && vl == ml && id >= 0) ? (unsigned)id : 0x7fffffffu);
    bv = __hiloint2double((int)mh, (int)ml);
    bid = (mid == 0x7fffffffu) ? -1 : (int)mid;
}

// Every warp finds its best candidate and publishes, next to the score, the Householder reflector of that candidate's column of y
// for the next pivot index jn (LAPACK dlarfg: beta = -sign(alpha) ||(alpha, x)||, tau = (beta - alpha) / beta, v = [1; x / (alpha -
// beta)]; rows < jn are dead) -- eight reflectors in parallel before the barrier instead of one derived by every thread behind it.
// After ONE barrier every warp picks the block winner from the eight entries.  KSMIN: k-steps below it are dead.
template <int KSMIN, bool GEN>
__device__ __forceinline__ ArgMax publish_and_pick(Smem& sm, const Filter& f, double s0, double s1, bool alive0, bool alive1, int pos0, int jn) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, q = lane & 3, pq = lane >> 2;
    double mv = 0.0; int mi = -1;
    if (alive0) { mv = s0; mi = pos0; }
    if (alive1 && (mi < 0 || s1 > mv)) { mv = s1; mi = pos0 + 8; }
    double wbv; int wp;
    warp_argmax_pos(mv, mi, wbv, wp);
    if (wp >= 0) {                               // warp-uniform
        double ys[8];
        if ((wp >> 3) & 1) {                     // warp-uniform too: which of the thread's two candidates
#pragma unroll
            for (int ks = KSMIN; ks < 8; ++ks) ys[ks] = f.Y[ks][1];
        } else {
#pragma unroll
            for (int ks = KSMIN; ks < 8; ++ks) ys[ks] = f.Y[ks][0];
        }
        // from a filter step jn lies in [4 KSMIN, 4 KSMIN + 4]: only the first two k-steps can hold dead rows or the pivot
        constexpr int KSEL = GEN ? 8 : (KSMIN + 2 < 8 ? KSMIN + 2 : 8);
        double e = 0.0, al = 0.0;
#pragma unroll
        for (int ks = KSMIN; ks < 8; ++ks) {
            if (ks < KSEL) {
                const int k = 4 * ks + q;
                if (k > jn) e = fma(ys[ks], ys[ks], e);
                if (k == jn) al = ys[ks];
            } else {
                e = fma(ys[ks], ys[ks], e);
            }
        }
        e = quad_sum(e); al = quad_sum(al);      // al: exact, one lane holds the pivot entry
        double tau = 0.0, sc = 0.0;
        if (e != 0.0) {
            const double nn = fma(al, al, e);
            const double beta = -copysign(nn * fast_rsqrt(nn), al);
            tau = (beta - al) * fast_rcp(beta);
            sc = fast_rcp(al - beta);
        }
        if (pq == (wp & 7)) {                    // the quad that owns the winner
            double* dst = sm.xw + warp * 32 + q;
#pragma unroll
            for (int ks = KSMIN; ks < 8; ++ks) {
                if (ks < KSEL) {
                    const int k = 4 * ks + q;
                    dst[4 * ks] = (k < jn) ? 0.0 : ((k == jn) ? 1.0 : ys[ks] * sc);
                } else {
                    dst[4 * ks] = ys[ks] * sc;
                }
            }
            if (q == 0) sm.tauw[warp] = tau;
        }
    }
    if (lane == 0) { sm.wv[warp] = wbv; sm.wi[warp] = wp; }
    __syncthreads();
    ArgMax r;
    warp_argmax_pos(lane < NWARP ? sm.wv[lane] : 0.0, lane < NWARP ? sm.wi[lane] : -1, r.v, r.id);
    return r;
}

// One accepted point: position `bpos` becomes pivot j = f.jY (k-step KSJ = j / 4).  Reflector on y and W (registers), column maxima,
// W D^-2 to shared memory, scores of the remaining candidates on the tensor path, next winner.  `last`: only W is updated.
template <int KSJ>
__device__ __forceinline__ ArgMax filter_step(Smem& sm, Filter& f, int n, int bpos, bool last, bool& alive0, bool& alive1, int pos0) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = tid >> 3, g = tid & 7;         // W role: row r, columns 4g..4g+3
    const int q = lane & 3, pq = lane >> 2;      // score role: k = 4 ks + q, candidates 16 warp + 8 ct + pq
    const int j = f.jY, bw = bpos >> 4;
    const double* vcol = sm.xw + bw * 32;
    const double tau = sm.tauw[bw];
    if (warp == bw && pq == (bpos & 7)) { if ((bpos >> 3) & 1) alive1 = false; else alive0 = false; }
    // ---- y <- H y (rows k < j are dead, row j dies now)
    if (!last && tau != 0.0) {
        double v[8];
        double g0 = 0.0, g1 = 0.0;
#pragma unroll
        for (int ks = KSJ; ks < 8; ++ks) { v[ks] = vcol[4 * ks + q]; g0 = fma(v[ks], f.Y[ks][0], g0); g1 = fma(v[ks], f.Y[ks][1], g1); }
        g0 = quad_sum(g0) * tau; g1 = quad_sum(g1) * tau;
#pragma unroll
        for (int ks = KSJ; ks < 8; ++ks) { f.Y[ks][0] = fma(-g0, v[ks], f.Y[ks][0]); f.Y[ks][1] = fma(-g1, v[ks], f.Y[ks][1]); }
    }
    // ---- W <- W H (row r, columns 4g..4g+3; column groups below KSJ are dead)
    if (tau != 0.0) {
        double vc[4] = {0.0, 0.0, 0.0, 0.0};
        if (g >= KSJ) {
            const double2* s2 = reinterpret_cast<const double2*>(vcol + 4 * g);
            const double2 a_ = s2[0], b_ = s2[1];
            vc[0] = a_.x; vc[1] = a_.y; vc[2] = b_.x; vc[3] = b_.y;
        }
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) a = fma(f.w[i], vc[i], a);
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        a *= tau;
#pragma unroll
        for (int i = 0; i < 4; ++i) f.w[i] = fma(-a, vc[i], f.w[i]);
    }
    f.jY = j + 1;
    ArgMax none; none.v = 0.0; none.id = -1;
    if (last) return none;
    const int jn = j + 1, ks0 = jn >> 2;         // ks0 is KSJ or KSJ + 1
    // ---- column maxima of |W|: over the four rows of this warp by shuffles, over the warps through shared memory
    if (g >= KSJ) {
        const unsigned gm = 0x01010101u << g;
        double m[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            m[i] = fabs(f.w[i]);
            m[i] = maxnn(m[i], __shfl_xor_sync(gm, m[i], 8));
            m[i] = maxnn(m[i], __shfl_xor_sync(gm, m[i], 16));
        }
        if (lane < 8) {
            double2* dst = reinterpret_cast<double2*>(sm.cm + warp * 32 + 4 * g);
            dst[0] = make_double2(m[0], m[1]); dst[1] = make_double2(m[2], m[3]);
        }
    }
    __syncthreads();
    // ---- B operand: W D^-2 on the live columns, zero elsewhere.  Lane c of every warp derives 1 / D_c^2, the row owners fetch
    // their four by shuffle.
    {
        double dinv = 0.0;
        if (lane >= 4 * KSJ) {
            double D = 0.0;
#pragma unroll
            for (int w = 0; w < NWARP; ++w) D = maxnn(D, sm.cm[w * 32 + lane]);
            if (lane >= jn && lane < n && D > 0.0) dinv = fast_rcp(D * D);
        }
        double o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = f.w[i] * __shfl_sync(0xffffffffu, dinv, 4 * g + i);
        if (g >= ks0) {
            double2* dst = reinterpret_cast<double2*>(sm.Wt + g * WTS + r * 4);
            dst[0] = make_double2(o[0], o[1]); dst[1] = make_double2(o[2], o[3]);
        }
    }
    __syncthreads();
    // ---- scores^T = y^T (W D^-2)^T on the FP64 tensor path; the thread ends up with rows 8 rt + 2 q + {0,1} of its two candidates
    double acc[2][4][2];
#pragma unroll
    for (int ct = 0; ct < 2; ++ct)
#pragma unroll
        for (int rt = 0; rt < 4; ++rt) { acc[ct][rt][0] = 0.0; acc[ct][rt][1] = 0.0; }
#pragma unroll
    for (int ks = KSJ; ks < 8; ++ks) {
        if (ks > KSJ || ks0 == KSJ) {
            const double* bp = sm.Wt + ks * WTS + lane;
            double bf[4];
#pragma unroll
            for (int rt = 0; rt < 4; ++rt) bf[rt] = bp[32 * rt];
#pragma unroll
            for (int rt = 0; rt < 4; ++rt) { dmma884(acc[0][rt], f.Y[ks][0], bf[rt]); dmma884(acc[1][rt], f.Y[ks][1], bf[rt]); }
        }
    }
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int rt = 0; rt < 4; ++rt) {
        s0 = maxnn(s0, maxnn(fabs(acc[0][rt][0]), fabs(acc[0][rt][1])));
        s1 = maxnn(s1, maxnn(fabs(acc[1][rt][0]), fabs(acc[1][rt][1])));
    }
    s0 = quad_max(s0); s1 = quad_max(s1);
    return publish_and_pick<KSJ, false>(sm, f, s0, s1, alive0, alive1, pos0, jn);
}

// One run of the filter over the candidates with (flags & want) == want and !(flags & (CF_USED | avoid)); picks are appended to
// out[] (1-based ids); returns their number.  W and jY continue from the previous run (AffinelyIndependentPoints.jl:14-42).
__device__ __forceinline__ int filter_run_mma(Smem& sm, Filter& f, const double* __restrict__ sites, int n, int n_db, unsigned want,
                                              unsigned avoid, double piv, int n_wanted, int* out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = tid >> 3, g = tid & 7;
    const int q = lane & 3, pq = lane >> 2;
    // ---- compact candidate list (one warp; ascending ids, so "first maximiser" = smallest position)
    if (warp == 0) {
        int base = 0;
        for (int i0 = 0; i0 < n_db; i0 += 32) {
            const int id = i0 + lane;
            bool act = false;
            if (id < n_db) { const unsigned fb = sm.fl[id]; act = ((fb & want) == want) && !(fb & (CF_USED | avoid)); }
            const unsigned msk = __ballot_sync(0xffffffffu, act);
            if (act) sm.cl[base + __popc(msk & ((1u << lane) - 1u))] = id;
            base += __popc(msk);
        }
        if (lane == 0) sm.ctl[1] = base;
    }
    if (f.jY > 0) {                              // W is no longer the identity: the coefficients need it from shared memory
        double2* dst = reinterpret_cast<double2*>(sm.Wt + r * 32 + 4 * g);
        dst[0] = make_double2(f.w[0], f.w[1]); dst[1] = make_double2(f.w[2], f.w[3]);
    }
    __syncthreads();
    const int nc = sm.ctl[1];
    if (nc == 0) return 0;
    // ---- y = W' s for this thread's two candidates, and || s ||_inf for the unconditional first pick (:51-69)
    const int pos0 = 16 * warp + pq;
    bool alive0 = pos0 < nc, alive1 = pos0 + 8 < nc;
    double s0 = 0.0, s1 = 0.0;
    {
        const double* sa = alive0 ? sites + (size_t)sm.cl[pos0] * n : nullptr;
        const double* sb = alive1 ? sites + (size_t)sm.cl[pos0 + 8] * n : nullptr;
        if (f.jY == 0) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const int k = 4 * ks + q;
                const bool in = k < n;
                const double xk = in ? sm.x[k] : 0.0;
                f.Y[ks][0] = (in && alive0) ? sa[k] - xk : 0.0;
                f.Y[ks][1] = (in && alive1) ? sb[k] - xk : 0.0;
                s0 = maxnn(s0, fabs(f.Y[ks][0])); s1 = maxnn(s1, fabs(f.Y[ks][1]));
            }
            s0 = quad_max(s0); s1 = quad_max(s1);
        } else {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) { f.Y[ks][0] = 0.0; f.Y[ks][1] = 0.0; }
            for (int i = 0; i < n; ++i) {
                const double xi = sm.x[i];
                const double a = alive0 ? sa[i] - xi : 0.0, b = alive1 ? sb[i] - xi : 0.0;
                s0 = maxnn(s0, fabs(a)); s1 = maxnn(s1, fabs(b));
                const double* wrow = sm.Wt + i * 32 + q;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) { const double wv_ = wrow[4 * ks]; f.Y[ks][0] = fma(wv_, a, f.Y[ks][0]); f.Y[ks][1] = fma(wv_, b, f.Y[ks][1]); }
            }
        }
    }
    ArgMax best = publish_and_pick<0, true>(sm, f, s0, s1, alive0, alive1, pos0, f.jY);
    int found = 0;
    for (;;) {
        if (best.id < 0) break;                                  // no candidate left
        if (found > 0 && !(best.v > piv)) break;                 // AffinelyIndependentPoints.jl:92 (the first pick is unconditional)
        if (tid == 0) { const int id = sm.cl[best.id]; sm.fl[id] |= CF_USED; out[found] = id + 1; }
        found += 1;
        const bool last = (found == n_wanted);
        switch (f.jY >> 2) {
        case 0: best = filter_step<0>(sm, f, n, best.id, last, alive0, alive1, pos0); break;
        case 1: best = filter_step<1>(sm, f, n, best.id, last, alive0, alive1, pos0); break;
        case 2: best = filter_step<2>(sm, f, n, best.id, last, alive0, alive1, pos0); break;
        case 3: best = filter_step<3>(sm, f, n, best.id, last, alive0, alive1, pos0); break;
        case 4: best = filter_step<4>(sm, f, n, best.id, last, alive0, alive1, pos0); break;
        case 5: best = filter_step<5>(sm, f, n, best.id, last, alive0, alive1, pos0); break;
        case 6: best = filter_step<6>(sm, f, n, best.id, last, alive0, alive1, pos0); break;
        default: best = filter_step<7>(sm, f, n, best.id, last, alive0, alive1, pos0); break;
        }
        if (last) break;
    }
    return found;
}

}  // namespace

__global__ void __launch_bounds__(NTHR, 2) select_rounds123_mma_kernel(SelectParams P) {
    __shared__ __align__(16) Smem sm;
    const int b = blockIdx.x, n = P.n, tid = threadIdx.x;
    const int r = tid >> 3, g = tid & 7;
    const int n_db = P.n_db[b];
    const double* sites = P.sites + (size_t)b * P.db_stride * n;
    const int x_index = P.x_index[b] - 1;
    const double delta = P.delta[b];
    const double delta_1 = P.cfg.theta_enlarge_1 * delta;
    const double piv = P.cfg.theta_pivot * delta_1;
    const double delta_2 = P.cfg.theta_enlarge_2 * P.delta_max;
    int* r1 = P.r1 + (size_t)b * n; int* r2 = P.r2 + (size_t)b * n;
    double* r3s = P.r3_sites + (size_t)b * n * n;
    double* dirs = P.dirs + (size_t)b * n * n;

    if (tid < 32) {
        const bool in = tid < n;
        const double xi = in ? P.x[(size_t)b * n + tid] : 0.0;
        sm.x[tid] = xi;
        const double gl = in ? P.glb[tid] : 0.0, gu = in ? P.gub[tid] : 0.0;
        sm.lb1[tid] = fmax(gl, xi - delta_1); sm.ub1[tid] = fmin(gu, xi + delta_1);     // utilities.jl:290-294
        sm.lb2[tid] = fmax(gl, xi - delta_2); sm.ub2[tid] = fmin(gu, xi + delta_2);
    }
    __syncthreads();
    // box scan (Databases.jl:324-327): inclusive bounds, both boxes at once; the box-2 bounds go to round 4
    // One warp per site, one lane per coordinate (n <= 32): a site is one coalesced row read, four rows in flight per warp, the
    // verdict two votes -- a thread per site walked its 240-byte row with 30 dependent, uncoalesced loads (14 % of the kernel's samples).
    {
        const int lane = tid & 31, warp = tid >> 5;
        const bool lin = lane < n;
        const double l1 = lin ? sm.lb1[lane] : 0.0, u1 = lin ? sm.ub1[lane] : 0.0, l2 = lin ? sm.lb2[lane] : 0.0, u2 = lin ? sm.ub2[lane] : 0.0;
        for (int id0 = warp; id0 < n_db; id0 += 4 * NWARP) {
            double v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const int id = id0 + q * NWARP; v[q] = (lin && id < n_db) ? sites[(size_t)id * n + lane] : 0.0; }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int id = id0 + q * NWARP;
                const bool in1 = __all_sync(0xffffffffu, !lin || (l1 <= v[q] && v[q] <= u1));
                const bool in2 = __all_sync(0xffffffffu, !lin || (l2 <= v[q] && v[q] <= u2));
                if (lane == 0 && id < n_db) sm.fl[id] = (unsigned char)((id != x_index) ? ((in1 ? CF_BOX1 : 0) | (in2 ? CF_BOX2 : 0)) : 0);
            }
        }
    }
    if (tid < n) { P.lb2[(size_t)b * n + tid] = sm.lb2[tid]; P.ub2[(size_t)b * n + tid] = sm.ub2[tid]; }

    bool ensure_fl = P.flags_in[2 * b] != 0;
    bool force_rebuild = P.flags_in[2 * b + 1] != 0;
    bool rebuilt = false;
    Filter f;
    int n_r1, n_r2, n_r3, n_dirs;
    bool fully_linear;
    for (;;) {   // at most two passes: the second is the coordinate rebuild (RbfModel.jl:634-637)
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) f.w[i] = (r == 4 * g + i && r < n) ? 1.0 : 0.0;
        f.jY = 0;
        for (int id = tid; id < n_db; id += NTHR) sm.fl[id] &= (unsigned char)~CF_USED;
        __syncthreads();
        n_r1 = n_r2 = n_r3 = 0; fully_linear = false;
        const bool skip_search = force_rebuild || !P.cfg.optimized_sampling;
        if (skip_search) {                       // RbfModel.jl:564-569: directions e_1..e_n in natural order
            for (int e = tid; e < n * n; e += NTHR) dirs[e] = ((e % n) == (e / n)) ? 1.0 : 0.0;
            n_dirs = n;
        } else {
            n_r1 = filter_run_mma(sm, f, sites, n, n_db, CF_BOX1, 0, piv, n, r1);
            // improving directions = reverse(eachcol(Z)), Z = live columns of W scaled by their inf-norm (RbfModel.jl:232)
            __syncthreads();
            {
                double m[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    m[i] = fabs(f.w[i]);
                    m[i] = fmax(m[i], __shfl_xor_sync(0xffffffffu, m[i], 8));
                    m[i] = fmax(m[i], __shfl_xor_sync(0xffffffffu, m[i], 16));
                }
                if ((tid & 31) < 8) {
                    double2* dst = reinterpret_cast<double2*>(sm.cm + (tid >> 5) * 32 + 4 * g);
                    dst[0] = make_double2(m[0], m[1]); dst[1] = make_double2(m[2], m[3]);
                }
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int c = 4 * g + i;
                double D = 0.0;
                for (int w = 0; w < NWARP; ++w) D = fmax(D, sm.cm[w * 32 + c]);
                if (c >= f.jY && c < n && r < n) dirs[r + (size_t)(n - 1 - c) * n] = f.w[i] / D;
            }
            n_dirs = n - f.jY;
        }
        int n_missing = n - n_r1;
        const bool approx = fabs(delta - P.delta_max) <= P.approx_rtol * fmax(fabs(delta), fabs(P.delta_max));
        if (n_missing == 0 || skip_search || ensure_fl || (approx && P.cfg.theta_enlarge_1 == P.cfg.theta_enlarge_2)) {
            fully_linear = true;                 // RbfModel.jl:588-591
        } else {                                 // round 2: box 2, excluding every round-1 candidate
            __syncthreads();
            n_r2 = filter_run_mma(sm, f, sites, n, n_db, CF_BOX2, CF_BOX1, piv, n_missing, r2);
        }
        n_missing -= n_r2;
        bool failed = false;
        if (n_missing > 0) {                     // round 3, RbfModel.jl:269-307
            int n_new = min(n_missing, P.max_new[b]); if (n_new < 0) n_new = 0;
            bool fl = n_new >= n_missing;
            __syncthreads();
            if (tid == 0) sm.ctl[0] = 0;
            __syncthreads();
            for (int i = tid; i < n_new; i += NTHR) {
                const double* d = dirs + (size_t)i * n;
                const double len = intersect_box_absmax(n, sm.x, d, sm.lb1, sm.ub1);
                double on = 0.0;
                for (int rr = 0; rr < n; ++rr) { const double o = len * d[rr]; r3s[(size_t)i * n + rr] = sm.x[rr] + o; on = fmax(on, fabs(o)); }
                if (on <= piv) atomicOr(&sm.ctl[0], 1);
            }
            __syncthreads();
            const bool any_small = sm.ctl[0] != 0;
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
    unsigned char* cflags = P.cflags + (size_t)b * P.db_stride;
    for (int id = tid; id < n_db; id += NTHR) cflags[id] = sm.fl[id];
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

bool select_mma_eligible(int n, int db_stride) { return n <= 32 && db_stride <= NCMAX; }

cudaError_t launch_select_rounds123_mma(const SelectParams& P, cudaStream_t s) {
    select_rounds123_mma_kernel<<<P.B, NTHR, 0, s>>>(P);
    return cudaGetLastError();
}

}  // namespace mrbf
