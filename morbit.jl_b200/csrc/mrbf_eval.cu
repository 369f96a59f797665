// Batched surrogate evaluation and Jacobians: eval_models / get_gradient / get_jacobian
// (src/models/RbfModel.jl:783-800 -> RBF model call, grad, jac; callers descent.jl:150-185, 196, algorithm.jl:766-767).
//
//   m_l(x)      = sum_i w_il phi(||x - c_i||) + lambda_0l + sum_r lambda_(r+1)l x_r
//   grad m_l(x) = sum_i w_il psi(||x - c_i||) (x - c_i) + lambda_(1..n)l ,   psi = phi'(rho)/rho
//
// Tiled kernel (n <= 64): a CTA owns 64 trial points and streams the centres through shared memory 64 at a
// time.  Both contractions are register-tiled FP64 GEMMs (4 x 4 micro-tiles, 16 independent DFMA chains):
//   phase 1   D = X' C'^T          -> rho^2 = |x'|^2 + |c'|^2 - 2 D   (coordinates centred at the first
//             centre, i.e. the trust-region centre, so the cancellation error is O(eps Delta^2))
//   phase 2   J_l -= (psi .* w_l) C'  through a shared 64 x 64 tile,   J_l += x' rowsum(psi .* w_l)
// Values are accumulated in registers and reduced over the 16 threads of a point row with shuffles.
// The generic kernel (any n) is the simple one-thread-per-point form used for n > 64.
#include "mrbf_common.cuh"
#include "mrbf_kernels.h"

namespace mrbf {

constexpr int TM = 64;    // trial points per CTA
constexpr int TN = 64;    // centres per tile
constexpr int GLD = TN + 1;

template <int CQ, bool WANT_J>
__global__ void __launch_bounds__(256, 1) eval_tile_kernel(EvalParams P, int l0, int kk) {
    constexpr int ND = 16 * CQ;           // padded coordinate count
    constexpr int KG = WANT_J ? 2 : 4;    // outputs handled per pass
    extern __shared__ __align__(16) double smem[];
    double* Xs = smem;                    // ND x TM   coordinate-major
    double* Cs = Xs + ND * TM;            // ND x TN   coordinate-major (phase 1)
    double* Cc = Cs + ND * TN;            // TN x ND   centre-major    (phase 2)
    double* Gs = Cc + TN * ND;            // TM x GLD
    double* xx = Gs + TM * GLD;           // TM
    double* cc = xx + TM;                 // TN
    double* Wt = cc + TN;                 // KG x TN
    double* xref = Wt + KG * TN;          // ND
    const int b = blockIdx.y, n = P.n, k = P.k, tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const long long m0 = (long long)blockIdx.x * TM;
    const int N = P.N[b];
    const double* centers = P.centers + (size_t)b * P.train_stride * n;
    const double* w = P.w + (size_t)b * P.train_stride * k;
    const int pl = P.p > 0 ? P.p : 1;
    const double* lam = P.lam + (size_t)b * pl * k;
    const double* X = P.X + (size_t)b * P.M * n;
    RadFn rf; rf.kernel = P.kernel; rf.ibeta = P.ibeta; rf.sgn = P.sgn; rf.alpha2 = P.alpha2[b];

    for (int c = tid; c < ND; c += 256) xref[c] = (c < n) ? centers[c] : 0.0;
    __syncthreads();
    for (int e = tid; e < TM * n; e += 256) {              // coalesced read of the point tile (AoS rows)
        const int pt = e / n, c = e % n;
        const long long mi = m0 + pt;
        Xs[c * TM + pt] = (mi < P.M) ? X[(size_t)mi * n + c] - xref[c] : 0.0;
    }
    for (int e = tid; e < TM * (ND - n); e += 256) { const int pt = e % TM, c = n + e / TM; Xs[c * TM + pt] = 0.0; }
    __syncthreads();
    if (tid < TM) { double s = 0.0; for (int c = 0; c < n; ++c) { double a = Xs[c * TM + tid]; s = fma(a, a, s); } xx[tid] = s; }

    double accY[4][KG];
    double accJ[WANT_J ? KG : 1][4][CQ];
    double gsum[WANT_J ? KG : 1][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int l = 0; l < KG; ++l) accY[a][l] = 0.0;
    if constexpr (WANT_J) {
#pragma unroll
        for (int l = 0; l < KG; ++l)
#pragma unroll
            for (int a = 0; a < 4; ++a) { gsum[l][a] = 0.0;
#pragma unroll
                for (int c = 0; c < CQ; ++c) accJ[l][a][c] = 0.0; }
    }

    for (int c0 = 0; c0 < N; c0 += TN) {
        __syncthreads();
        for (int e = tid; e < TN * n; e += 256) {           // centre tile, both layouts
            const int j = e / n, c = e % n;
            const double val = (c0 + j < N) ? centers[(size_t)(c0 + j) * n + c] - xref[c] : 0.0;
            Cs[c * TN + j] = val; Cc[j * ND + c] = val;
        }
        for (int e = tid; e < TN * (ND - n); e += 256) { const int j = e / (ND - n), c = n + e % (ND - n); Cc[j * ND + c] = 0.0; }
        for (int e = tid; e < KG * TN; e += 256) {
            const int l = e / TN, j = e % TN;
            Wt[e] = (l < kk && c0 + j < N) ? w[(size_t)(c0 + j) * k + l0 + l] : 0.0;
        }
        __syncthreads();
        if (tid < TN) { double s = 0.0; for (int c = 0; c < n; ++c) { double a = Cs[c * TN + tid]; s = fma(a, a, s); } cc[tid] = s; }
        // ---- phase 1: 4 x 4 dot products
        double d[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < 4; ++j) d[a][j] = 0.0;
#pragma unroll 2
        for (int c = 0; c < n; ++c) {
            const double2 x01 = *reinterpret_cast<const double2*>(Xs + c * TM + 4 * ty);
            const double2 x23 = *reinterpret_cast<const double2*>(Xs + c * TM + 4 * ty + 2);
            const double2 c01 = *reinterpret_cast<const double2*>(Cs + c * TN + 4 * tx);
            const double2 c23 = *reinterpret_cast<const double2*>(Cs + c * TN + 4 * tx + 2);
            const double xv[4] = {x01.x, x01.y, x23.x, x23.y};
            const double cv[4] = {c01.x, c01.y, c23.x, c23.y};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int j = 0; j < 4; ++j) d[a][j] = fma(xv[a], cv[j], d[a][j]);
        }
        __syncthreads();                                     // cc ready
        double psi[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double r2 = fma(-2.0, d[a][j], xx[4 * ty + a] + cc[4 * tx + j]);
                r2 = fmax(r2, 0.0);
                double ph, ps;
                if constexpr (WANT_J) rad_phi_psi(rf, r2, ph, ps); else { ph = rad_phi(rf, r2); ps = 0.0; }
                d[a][j] = ph; psi[a][j] = ps;
            }
#pragma unroll
        for (int l = 0; l < KG; ++l)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double wj = Wt[l * TN + 4 * tx + j];
#pragma unroll
                for (int a = 0; a < 4; ++a) accY[a][l] = fma(d[a][j], wj, accY[a][l]);
            }
        if constexpr (WANT_J) {
            for (int l = 0; l < kk; ++l) {
                __syncthreads();                             // previous pass finished reading Gs
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int j = 0; j < 4; ++j) Gs[(4 * ty + a) * GLD + 4 * tx + j] = psi[a][j] * Wt[l * TN + 4 * tx + j];
                __syncthreads();
                // ---- phase 2: accJ[l] += G (64 x 64) * C' (64 x ND), micro-tile 4 points x CQ coordinates
#pragma unroll 4
                for (int j = 0; j < TN; ++j) {
                    double g[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) g[a] = Gs[(4 * ty + a) * GLD + j];
                    double cv[CQ];
#pragma unroll
                    for (int c = 0; c < CQ; ++c) cv[c] = Cc[j * ND + tx * CQ + c];
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        if (l == 0) gsum[0][a] += g[a]; else gsum[KG - 1][a] += g[a];
#pragma unroll
                        for (int c = 0; c < CQ; ++c) {
                            if (l == 0) accJ[0][a][c] = fma(g[a], cv[c], accJ[0][a][c]);
                            else accJ[KG - 1][a][c] = fma(g[a], cv[c], accJ[KG - 1][a][c]);
                        }
                    }
                }
            }
        }
    }
    // ---- epilogue
    // values: reduce the 16 partial sums of a point row (lanes tx = 0..15 of a half warp)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int l = 0; l < KG; ++l) {
            double s = accY[a][l];
            s += __shfl_xor_sync(0xffffffffu, s, 8); s += __shfl_xor_sync(0xffffffffu, s, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 1);
            accY[a][l] = s;
        }
    if (P.Y && tx == 0) {
        for (int a = 0; a < 4; ++a) {
            const long long mi = m0 + 4 * ty + a;
            if (mi >= P.M) continue;
            for (int l = 0; l < kk; ++l) {
                double s = accY[a][l];
                if (P.deg >= 0) s += lam[l0 + l];
                if (P.deg >= 1) {
                    double t = 0.0;
                    for (int c = 0; c < n; ++c) t = fma(lam[(size_t)(c + 1) * k + l0 + l], Xs[c * TM + 4 * ty + a] + xref[c], t);
                    s += t;
                }
                P.Y[((size_t)b * P.M + mi) * k + l0 + l] = s;
            }
        }
    }
    if constexpr (WANT_J) if (P.J) {
        for (int l = 0; l < kk; ++l)
            for (int a = 0; a < 4; ++a) {
                const long long mi = m0 + 4 * ty + a;
                if (mi >= P.M) continue;
                double* Jrow = P.J + (((size_t)b * P.M + mi) * k + l0 + l) * n;
#pragma unroll
                for (int c = 0; c < CQ; ++c) {
                    const int r = tx * CQ + c;
                    if (r < n) {
                        const double gs = (l == 0) ? gsum[0][a] : gsum[KG - 1][a];
                        const double aj = (l == 0) ? accJ[0][a][c] : accJ[KG - 1][a][c];
                        double val = fma(Xs[r * TM + 4 * ty + a], gs, -aj);
                        if (P.deg >= 1) val += lam[(size_t)(r + 1) * k + l0 + l];
                        Jrow[r] = val;
                    }
                }
            }
    }
}

// Values for many trial points when n > 64 (BASELINE config C4: n = 200 -- the Armijo batches of descent.jl:150-185): the tile of
// 64 trial points stays in shared memory with ALL its coordinates (coordinate-major, centred at the first centre), the centre
// tiles are streamed through shared memory in chunks of WC coordinates, and the distance contraction runs on the same 4 x 4
// register micro-tiles as eval_tile_kernel.  The generic one-thread-per-point kernel it replaces ran at ~1 % of the FP64 peak.
constexpr int WC = 32;          // coordinates per streamed chunk
constexpr int CLD = TN + 2;     // padded row of the centre chunk (16-byte aligned rows, 4-way instead of 32-way store conflicts)

__global__ void __launch_bounds__(256, 1) eval_wide_kernel(EvalParams P, int l0, int kk, int npad) {
    constexpr int KG = 4;
    extern __shared__ __align__(16) double smem[];
    double* Xs = smem;                    // npad x TM
    double* Cs = Xs + (size_t)npad * TM;  // WC x CLD
    double* xx = Cs + WC * CLD;           // TM
    double* cc = xx + TM;                 // TN
    double* Wt = cc + TN;                 // KG x TN
    double* xref = Wt + KG * TN;          // npad
    const int b = blockIdx.y, n = P.n, k = P.k, tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const long long m0 = (long long)blockIdx.x * TM;
    const int N = P.N[b];
    const double* centers = P.centers + (size_t)b * P.train_stride * n;
    const double* w = P.w + (size_t)b * P.train_stride * k;
    const int pl = P.p > 0 ? P.p : 1;
    const double* lam = P.lam + (size_t)b * pl * k;
    const double* X = P.X + (size_t)b * P.M * n;
    RadFn rf; rf.kernel = P.kernel; rf.ibeta = P.ibeta; rf.sgn = P.sgn; rf.alpha2 = P.alpha2[b];

    for (int c = tid; c < npad; c += 256) xref[c] = (c < n) ? centers[c] : 0.0;
    __syncthreads();
    for (int e = tid; e < TM * n; e += 256) {              // coalesced read of the point tile (AoS rows)
        const int pt = e / n, c = e % n;
        const long long mi = m0 + pt;
        Xs[c * TM + pt] = (mi < P.M) ? X[(size_t)mi * n + c] - xref[c] : 0.0;
    }
    for (int e = tid; e < TM * (npad - n); e += 256) { const int pt = e % TM, c = n + e / TM; Xs[c * TM + pt] = 0.0; }
    __syncthreads();
    if (tid < TM) { double s = 0.0; for (int c = 0; c < n; ++c) { double a = Xs[c * TM + tid]; s = fma(a, a, s); } xx[tid] = s; }

    double accY[4][KG];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int l = 0; l < KG; ++l) accY[a][l] = 0.0;

    for (int c0 = 0; c0 < N; c0 += TN) {
        double d[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < 4; ++j) d[a][j] = 0.0;
        double ccp = 0.0;
        for (int k0 = 0; k0 < npad; k0 += WC) {
            __syncthreads();                                 // the previous chunk (and the previous tile's Wt / cc) has been consumed
            for (int e = tid; e < TN * WC; e += 256) {       // lanes over the chunk's coordinates: contiguous 256-byte reads per centre
                const int j = e / WC, c = e % WC;
                Cs[c * CLD + j] = (c0 + j < N && k0 + c < n) ? centers[(size_t)(c0 + j) * n + k0 + c] - xref[k0 + c] : 0.0;
            }
            if (k0 == 0)
                for (int e = tid; e < KG * TN; e += 256) {
                    const int l = e / TN, j = e % TN;
                    Wt[e] = (l < kk && c0 + j < N) ? w[(size_t)(c0 + j) * k + l0 + l] : 0.0;
                }
            __syncthreads();
            if (tid < TN) {
#pragma unroll 8
                for (int c = 0; c < WC; ++c) { const double a = Cs[c * CLD + tid]; ccp = fma(a, a, ccp); }
            }
            const double* xrow = Xs + (size_t)k0 * TM + 4 * ty;
#pragma unroll 4
            for (int c = 0; c < WC; ++c) {
                const double2 x01 = *reinterpret_cast<const double2*>(xrow + c * TM);
                const double2 x23 = *reinterpret_cast<const double2*>(xrow + c * TM + 2);
                const double2 c01 = *reinterpret_cast<const double2*>(Cs + c * CLD + 4 * tx);
                const double2 c23 = *reinterpret_cast<const double2*>(Cs + c * CLD + 4 * tx + 2);
                const double xv[4] = {x01.x, x01.y, x23.x, x23.y};
                const double cv[4] = {c01.x, c01.y, c23.x, c23.y};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int j = 0; j < 4; ++j) d[a][j] = fma(xv[a], cv[j], d[a][j]);
            }
        }
        if (tid < TN) cc[tid] = ccp;
        __syncthreads();
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double r2 = fma(-2.0, d[a][j], xx[4 * ty + a] + cc[4 * tx + j]);
                r2 = fmax(r2, 0.0);
                d[a][j] = rad_phi(rf, r2);
            }
#pragma unroll
        for (int l = 0; l < KG; ++l)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double wj = Wt[l * TN + 4 * tx + j];
#pragma unroll
                for (int a = 0; a < 4; ++a) accY[a][l] = fma(d[a][j], wj, accY[a][l]);
            }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int l = 0; l < KG; ++l) {
            double s = accY[a][l];
            s += __shfl_xor_sync(0xffffffffu, s, 8); s += __shfl_xor_sync(0xffffffffu, s, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 1);
            accY[a][l] = s;
        }
    if (P.Y && tx == 0) {
        for (int a = 0; a < 4; ++a) {
            const long long mi = m0 + 4 * ty + a;
            if (mi >= P.M) continue;
            for (int l = 0; l < kk; ++l) {
                double s = accY[a][l];
                if (P.deg >= 0) s += lam[l0 + l];
                if (P.deg >= 1) {
                    double t = 0.0;
                    for (int c = 0; c < n; ++c) t = fma(lam[(size_t)(c + 1) * k + l0 + l], Xs[c * TM + 4 * ty + a] + xref[c], t);
                    s += t;
                }
                P.Y[((size_t)b * P.M + mi) * k + l0 + l] = s;
            }
        }
    }
}

// Few trial points per instance (the Jacobian at the iterate, the rho test: M = 1; descent.jl:196, algorithm.jl:766): the tiled
// kernels would run 64-point tiles with one live row.  Here ONE WARP takes one (instance, point): lanes over the centres for
// phi / psi (each lane walks its centre's contiguous coordinates), the weighted psi go through shared memory, then lanes over
// the coordinates accumulate the Jacobian rows with coalesced reads of the centre matrix.
__global__ void __launch_bounds__(128) eval_small_kernel(EvalParams P, int warp_doubles) {
    extern __shared__ double sm_small[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long q = (long long)blockIdx.x * 4 + wib;
    if (q >= (long long)P.B * P.M) return;
    const int b = (int)(q / P.M);
    const long long mi = q % P.M;
    const int n = P.n, k = P.k;
    double* xs = sm_small + (size_t)wib * warp_doubles;      // n
    double* sw = xs + n;                                      // N x k   w_il * psi_i
    const int N = P.N[b];
    const double* centers = P.centers + (size_t)b * P.train_stride * n;
    const double* w = P.w + (size_t)b * P.train_stride * k;
    const int pl = P.p > 0 ? P.p : 1;
    const double* lam = P.lam + (size_t)b * pl * k;
    const double* x = P.X + ((size_t)b * P.M + mi) * n;
    RadFn rf; rf.kernel = P.kernel; rf.ibeta = P.ibeta; rf.sgn = P.sgn; rf.alpha2 = P.alpha2[b];
    for (int c = lane; c < n; c += 32) xs[c] = x[c];
    __syncwarp();
    const bool want_j = P.J != nullptr;
    for (int l0 = 0; l0 < k; l0 += 4) {
        const int kk = min(4, k - l0);
        double y[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = lane; i < N; i += 32) {
            const double* c = centers + (size_t)i * n;
            double r0 = 0.0, r1 = 0.0;
            int r = 0;
            for (; r + 1 < n; r += 2) { const double d0 = xs[r] - c[r], d1 = xs[r + 1] - c[r + 1]; r0 = fma(d0, d0, r0); r1 = fma(d1, d1, r1); }
            if (r < n) { const double d0 = xs[r] - c[r]; r0 = fma(d0, d0, r0); }
            double ph, ps;
            rad_phi_psi(rf, r0 + r1, ph, ps);
            for (int l = 0; l < kk; ++l) {
                const double wl = w[(size_t)i * k + l0 + l];
                y[l] = fma(wl, ph, y[l]);
                if (want_j) sw[i * k + l0 + l] = wl * ps;
            }
        }
        for (int l = 0; l < kk; ++l) {
            double sacc = warp_sum(y[l]);
            if (P.deg >= 0) sacc += lam[l0 + l];
            if (P.deg >= 1) {
                double t = 0.0;
                for (int r = lane; r < n; r += 32) t = fma(lam[(size_t)(r + 1) * k + l0 + l], xs[r], t);
                sacc += warp_sum(t);
            }
            if (P.Y && lane == 0) P.Y[((size_t)b * P.M + mi) * k + l0 + l] = sacc;
        }
    }
    if (!want_j) return;
    __syncwarp();
    double* J = P.J + ((size_t)b * P.M + mi) * k * n;
    for (int c = lane; c < n; c += 32) {
        const double xc = xs[c];
        for (int l0 = 0; l0 < k; l0 += 4) {
            const int kk = min(4, k - l0);
            double a[4] = {0.0, 0.0, 0.0, 0.0};
            for (int i = 0; i < N; ++i) {
                const double dc = xc - centers[(size_t)i * n + c];
                for (int l = 0; l < kk; ++l) a[l] = fma(sw[i * k + l0 + l], dc, a[l]);
            }
            for (int l = 0; l < kk; ++l) J[(size_t)(l0 + l) * n + c] = a[l] + ((P.deg >= 1) ? lam[(size_t)(c + 1) * k + l0 + l] : 0.0);
        }
    }
}

// Generic kernel: one thread per trial point, any n; outputs processed four at a time.
__global__ void __launch_bounds__(128) eval_generic_kernel(EvalParams P) {
    const int b = blockIdx.y, n = P.n, k = P.k;
    const long long mi = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (mi >= P.M) return;
    const int N = P.N[b];
    const double* centers = P.centers + (size_t)b * P.train_stride * n;
    const double* w = P.w + (size_t)b * P.train_stride * k;
    const int pl = P.p > 0 ? P.p : 1;
    const double* lam = P.lam + (size_t)b * pl * k;
    const double* x = P.X + ((size_t)b * P.M + mi) * n;
    RadFn rf; rf.kernel = P.kernel; rf.ibeta = P.ibeta; rf.sgn = P.sgn; rf.alpha2 = P.alpha2[b];
    double* Y = P.Y ? P.Y + ((size_t)b * P.M + mi) * k : nullptr;
    double* J = P.J ? P.J + ((size_t)b * P.M + mi) * k * n : nullptr;
    if (J) for (int e = 0; e < k * n; ++e) J[e] = 0.0;
    for (int l0 = 0; l0 < k; l0 += 4) {
        const int kk = min(4, k - l0);
        double y[4] = {0.0, 0.0, 0.0, 0.0};
        for (int i = 0; i < N; ++i) {
            const double* c = centers + (size_t)i * n;
            double r2 = 0.0;
            for (int r = 0; r < n; ++r) { double dd = x[r] - c[r]; r2 = fma(dd, dd, r2); }
            double ph, ps;
            rad_phi_psi(rf, r2, ph, ps);
            for (int l = 0; l < kk; ++l) {
                const double wl = w[(size_t)i * k + l0 + l];
                y[l] = fma(wl, ph, y[l]);
                if (J) {
                    const double g = wl * ps;
                    double* Jl = J + (size_t)(l0 + l) * n;
                    for (int r = 0; r < n; ++r) Jl[r] = fma(g, x[r] - c[r], Jl[r]);
                }
            }
        }
        for (int l = 0; l < kk; ++l) {
            double s = y[l];
            if (P.deg >= 0) s += lam[l0 + l];
            if (P.deg >= 1) {
                double t = 0.0;
                for (int r = 0; r < n; ++r) t = fma(lam[(size_t)(r + 1) * k + l0 + l], x[r], t);
                s += t;
                if (J) { double* Jl = J + (size_t)(l0 + l) * n; for (int r = 0; r < n; ++r) Jl[r] += lam[(size_t)(r + 1) * k + l0 + l]; }
            }
            if (P.deg >= 2) {                      // quadratic monomials x_a x_b, a <= b (a outer), behind the linear ones
                int q = n + 1;
                double* Jl = J ? J + (size_t)(l0 + l) * n : nullptr;
                for (int a = 0; a < n; ++a)
                    for (int bb = a; bb < n; ++bb, ++q) {
                        const double lq = lam[(size_t)q * k + l0 + l];
                        s = fma(lq, x[a] * x[bb], s);
                        if (Jl) { Jl[a] = fma(lq, x[bb], Jl[a]); Jl[bb] = fma(lq, x[a], Jl[bb]); }
                    }
            }
            if (Y) Y[l0 + l] = s;
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Values-only sweep on the FP64 tensor path (DMMA, mma.sync m8n8k4 f64) with TMA-staged centre tiles.
//
// The model is re-tiled once (eval_pack_kernel): per 64 centres one contiguous block
//     [64 x s centred centre rows | 64 squared norms | k x 64 coefficients],   s = padded row stride (= 4 mod 16 doubles,
// so the 8 rows x 4 doubles of an operand fragment fall into two conflict-free shared-memory wavefronts).
// A CTA owns 128 trial points (X' tile staged once) and streams the centre tiles with cp.async.bulk (TMA 1-D bulk
// copy) into a two-deep ring guarded by mbarriers, so the copy of tile t+1 overlaps the tensor work of tile t.
// 8 warps = 4 (point direction) x 2 (centre direction); a warp computes a 32 x 32 block of X'C'^T as 4 x 4 DMMA
// fragments (16 independent accumulator chains), then rho^2 = |x'|^2 + |c'|^2 - 2 D, phi, and the weighted row sums.
// ------------------------------------------------------------------------------------------------
constexpr int DM_TM = 128;   // trial points per CTA
constexpr int DM_TN = 64;    // centres per tile

__host__ __device__ inline int pack_stride(int n) {
    int s = (n + 3) & ~3;
    while ((s & 15) != 4 && (s & 15) != 12) s += 4;
    return s;
}
int eval_split_factor(long long M, int B, int pack_nt) {
    // CTAs of the tensor-path sweep: ceil(M / 128) x B, one per SM.  Below one wave the centre tiles are split so that ~one wave runs.
    const long long ctas = ((M + DM_TM - 1) / DM_TM) * (long long)B;
    if (ctas >= 148 || pack_nt < 2) return 1;
    long long z = 148 / ctas;                       // never more than one wave: a second, nearly empty wave costs more than it gains
    if (z < 1) z = 1;
    if (z > pack_nt) z = pack_nt;
    return (int)z;
}

int eval_pack_stride(int n) { return pack_stride(n); }

__global__ void eval_pack_kernel(PackParams P) {
    const int b = blockIdx.y, t = blockIdx.x, n = P.n, k = P.k, s = P.s, tid = threadIdx.x, nt = blockDim.x;
    const int N = P.N[b];
    const double* centers = P.centers + (size_t)b * P.train_stride * n;
    const double* w = P.w + (size_t)b * P.train_stride * k;
    double* out = P.pack + ((size_t)b * P.nt + t) * P.tile_doubles;
    double* cc = out + (size_t)DM_TN * s;
    double* Wt = cc + DM_TN;
    for (int e = tid; e < DM_TN * s; e += nt) {
        const int j = e / s, c = e % s, gi = t * DM_TN + j;
        out[e] = (gi < N && c < n) ? centers[(size_t)gi * n + c] - centers[c] : 0.0;
    }
    for (int j = tid; j < DM_TN; j += nt) {
        const int gi = t * DM_TN + j;
        double a = 0.0;
        if (gi < N) for (int c = 0; c < n; ++c) { double d = centers[(size_t)gi * n + c] - centers[c]; a = fma(d, d, a); }
        cc[j] = a;
    }
    for (int e = tid; e < k * DM_TN; e += nt) {
        const int l = e / DM_TN, j = e % DM_TN, gi = t * DM_TN + j;
        Wt[e] = (gi < N) ? w[(size_t)gi * k + l] : 0.0;
    }
}

cudaError_t launch_eval_pack(const PackParams& P, cudaStream_t s) {
    dim3 grid((unsigned)P.nt, (unsigned)P.B);
    eval_pack_kernel<<<grid, 128, 0, s>>>(P);
    return cudaGetLastError();
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

template <int RK, int KK>
__global__ void __launch_bounds__(256, 1) eval_dmma_kernel(EvalParams P, int l0) {
    constexpr int kk = KK;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int b = blockIdx.y, n = P.n, k = P.k, s = P.pack_s, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;
    const int tile_d = (int)P.pack_tile_doubles;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw);          // 2 mbarriers
    double* buf0 = reinterpret_cast<double*>(smem_raw + 128);
    double* buf1 = buf0 + tile_d;
    double* Xs = buf1 + tile_d;                  // DM_TM x s
    double* xx = Xs + DM_TM * s;                 // DM_TM
    double* Yp = xx + DM_TM;                     // 2 x DM_TM x 4  partial sums of the two centre-direction warps
    const long long m0 = (long long)blockIdx.x * DM_TM;
    const int N = P.N[b];
    const int ntiles = (N + DM_TN - 1) / DM_TN;
    const int Z = (int)gridDim.z, zz = (int)blockIdx.z;                                   // centre tiles [t_lo, t_hi) of this CTA
    const int t_lo = (int)(((long long)ntiles * zz) / Z), t_hi = (int)(((long long)ntiles * (zz + 1)) / Z);
    const double* centers = P.centers + (size_t)b * P.train_stride * n;
    const double* X = P.X + (size_t)b * P.M * n;
    const double* pack = P.pack + (size_t)b * P.pack_nt * P.pack_tile_doubles;
    const int pl = P.p > 0 ? P.p : 1;
    const double* lam = P.lam + (size_t)b * pl * k;
    RadFn rf; rf.kernel = P.kernel; rf.ibeta = P.ibeta; rf.sgn = P.sgn; rf.alpha2 = P.alpha2[b];
    const unsigned tile_bytes = (unsigned)(tile_d * sizeof(double));

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar[0])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar[1])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (t_hi > t_lo) {                       // first tile -> buffer 0
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar[0])), "r"(tile_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(buf0)), "l"(pack + (size_t)t_lo * tile_d), "r"(tile_bytes), "r"(smem_u32(&bar[0])) : "memory");
        }
    }
    // X' tile: rows centred at the first centre, zero padded to the row stride
    for (int e = tid; e < DM_TM * s; e += 256) {
        const int pt = e / s, c = e % s;
        const long long mi = m0 + pt;
        Xs[e] = (mi < P.M && c < n) ? X[(size_t)mi * n + c] - centers[c] : 0.0;
    }
    __syncthreads();
    if (tid < DM_TM) { double a = 0.0; for (int c = 0; c < n; ++c) { double v = Xs[tid * s + c]; a = fma(v, v, a); } xx[tid] = a; }
    __syncthreads();

    const int qr = lane >> 2, qc = lane & 3;     // fragment coordinates: row T/4, k-offset / column pair T%4
    double xr[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) xr[a] = xx[32 * wm + 8 * a + qr];
    double ysum[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int l = 0; l < 4; ++l) ysum[a][l] = 0.0;
    const int ksteps = ((n + 3) & ~3) >> 2;
    const double* xa = Xs + (32 * wm + qr) * s + qc;

    for (int t = t_lo; t < t_hi; ++t) {
        const int it = t - t_lo, cur = it & 1;
        if (tid == 0 && t + 1 < t_hi) {          // prefetch tile t+1 into the other buffer (its readers passed the barrier below)
            const int nb = cur ^ 1;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar[nb])), "r"(tile_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(nb ? buf1 : buf0)), "l"(pack + (size_t)(t + 1) * tile_d), "r"(tile_bytes), "r"(smem_u32(&bar[nb])) : "memory");
        }
        {                                        // wait for tile t
            const unsigned parity = (unsigned)((it >> 1) & 1);
            unsigned ok = 0;
            while (!ok) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&bar[cur])), "r"(parity) : "memory");
            }
        }
        const double* Cs = cur ? buf1 : buf0;
        const double* ccs = Cs + DM_TN * s;
        const double* Wt = ccs + DM_TN;
        double acc[4][4][2];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) { acc[a][c][0] = 0.0; acc[a][c][1] = 0.0; }
        const double* cb = Cs + (32 * wn + qr) * s + qc;
#pragma unroll 2
        for (int ks = 0; ks < ksteps; ++ks) {
            double fa[4], fb[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) fa[a] = xa[8 * a * s + 4 * ks];
#pragma unroll
            for (int c = 0; c < 4; ++c) fb[c] = cb[8 * c * s + 4 * ks];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) dmma884(acc[a][c], fa[a], fb[c]);
        }
        // epilogue of the tile: D[row qr + 8a][col 2 qc + e + 8c]
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = 32 * wn + 8 * c + 2 * qc + e;
                const double ccv = ccs[col];
                double wv[4];
#pragma unroll
                for (int l = 0; l < 4; ++l) wv[l] = (l < kk) ? Wt[l * DM_TN + col] : 0.0;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    double r2 = fma(-2.0, acc[a][c][e], xr[a] + ccv);
                    r2 = fmax(r2, 0.0);
                    const double ph = rad_phi_t<RK>(rf, r2);
#pragma unroll
                    for (int l = 0; l < 4; ++l) if (l < kk) ysum[a][l] = fma(ph, wv[l], ysum[a][l]);
                }
            }
        __syncthreads();                         // everyone is done with buffer `cur`
    }
    // reduce over the 4 lanes of a quad (column pairs), then over the two centre-direction warps through shared memory
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            double v = ysum[a][l];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (qc == 0) Yp[(wn * DM_TM + 32 * wm + 8 * a + qr) * 4 + l] = v;
        }
    __syncthreads();
    if (tid < DM_TM) {
        const long long mi = m0 + tid;
        if (mi < P.M) {
            for (int l = 0; l < kk; ++l) {
                double v = Yp[tid * 4 + l] + Yp[(DM_TM + tid) * 4 + l];
                if (Z > 1) { P.partY[(((size_t)zz * P.B + b) * P.M + mi) * k + l0 + l] = v; continue; }     // tail added by the reduction
                if (P.deg >= 0) v += lam[l0 + l];
                if (P.deg >= 1) {
                    double tsum = 0.0;
                    for (int c = 0; c < n; ++c) tsum = fma(lam[(size_t)(c + 1) * k + l0 + l], Xs[tid * s + c] + centers[c], tsum);
                    v += tsum;
                }
                P.Y[((size_t)b * P.M + mi) * k + l0 + l] = v;
            }
        }
    }
}

template <int RK, int KK>
static cudaError_t launch_dmma_t(const EvalParams& P, cudaStream_t s, int l0, size_t smem, dim3 grid) {
    cudaError_t e = raise_dyn_smem(eval_dmma_kernel<RK, KK>, smem);
    if (e != cudaSuccess) return e;
    eval_dmma_kernel<RK, KK><<<grid, 256, smem, s>>>(P, l0);
    return cudaGetLastError();
}
template <int RK>
static cudaError_t launch_dmma_k(const EvalParams& P, cudaStream_t s, int l0, int kk, size_t smem, dim3 grid) {
    switch (kk) {
    case 1: return launch_dmma_t<RK, 1>(P, s, l0, smem, grid);
    case 2: return launch_dmma_t<RK, 2>(P, s, l0, smem, grid);
    case 3: return launch_dmma_t<RK, 3>(P, s, l0, smem, grid);
    default: return launch_dmma_t<RK, 4>(P, s, l0, smem, grid);
    }
}
static cudaError_t launch_dmma(const EvalParams& P, cudaStream_t s, int* n_launches) {
    const int st = P.pack_s;
    const size_t smem = 128 + sizeof(double) * (2 * P.pack_tile_doubles + (size_t)DM_TM * st + DM_TM + 2 * DM_TM * 4);
    const long long tiles = (P.M + DM_TM - 1) / DM_TM;
    const int rk = rad_kind(P.kernel, P.ibeta);
    for (int l0 = 0; l0 < P.k; l0 += 4) {
        const int kk = (P.k - l0) < 4 ? (P.k - l0) : 4;
        dim3 grid((unsigned)tiles, (unsigned)P.B, (unsigned)(P.zsplit > 1 ? P.zsplit : 1));
        cudaError_t e;
        switch (rk) {
        case RK_CUBIC3: e = launch_dmma_k<RK_CUBIC3>(P, s, l0, kk, smem, grid); break;
        case RK_MQ: e = launch_dmma_k<RK_MQ>(P, s, l0, kk, smem, grid); break;
        case RK_GAUSS: e = launch_dmma_k<RK_GAUSS>(P, s, l0, kk, smem, grid); break;
        default: e = launch_dmma_k<RK_GENERIC>(P, s, l0, kk, smem, grid); break;
        }
        if (n_launches) ++*n_launches;
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// KO outputs per pass (1 or 2): the distance contraction and the radial functions are shared, only the weighting of psi and the
// second contraction are per output.
template <int RK, int KO>
__global__ void __launch_bounds__(256, 1) eval_dmma_jac_kernel(EvalParams P, int l0) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int b = blockIdx.y, n = P.n, k = P.k, s = P.pack_s, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_d = (int)P.pack_tile_doubles;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw);
    double* buf0 = reinterpret_cast<double*>(smem_raw + 128);
    double* buf1 = buf0 + tile_d;
    double* Xs = buf1 + tile_d;                  // DM_TM x s
    double* xx = Xs + DM_TM * s;                 // DM_TM
    const long long m0 = (long long)blockIdx.x * DM_TM;
    const int N = P.N[b];
    const int ntiles = (N + DM_TN - 1) / DM_TN;
    const int Z = (int)gridDim.z, zz = (int)blockIdx.z;                                   // centre tiles [t_lo, t_hi) of this CTA
    const int t_lo = (int)(((long long)ntiles * zz) / Z), t_hi = (int)(((long long)ntiles * (zz + 1)) / Z);
    const double* centers = P.centers + (size_t)b * P.train_stride * n;
    const double* X = P.X + (size_t)b * P.M * n;
    const double* pack = P.pack + (size_t)b * P.pack_nt * P.pack_tile_doubles;
    const int pl = P.p > 0 ? P.p : 1;
    const double* lam = P.lam + (size_t)b * pl * k;
    RadFn rf; rf.kernel = P.kernel; rf.ibeta = P.ibeta; rf.sgn = P.sgn; rf.alpha2 = P.alpha2[b];
    const unsigned tile_bytes = (unsigned)(tile_d * sizeof(double));
    const int ncb = (n + 7) >> 3;                // coordinate blocks of 8

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar[0])), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar[1])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (t_hi > t_lo) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar[0])), "r"(tile_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(buf0)), "l"(pack + (size_t)t_lo * tile_d), "r"(tile_bytes), "r"(smem_u32(&bar[0])) : "memory");
        }
    }
    for (int e = tid; e < DM_TM * s; e += 256) {
        const int pt = e / s, c = e % s;
        const long long mi = m0 + pt;
        Xs[e] = (mi < P.M && c < n) ? X[(size_t)mi * n + c] - centers[c] : 0.0;
    }
    __syncthreads();
    if (tid < DM_TM) { double a = 0.0; for (int c = 0; c < n; ++c) { double v = Xs[tid * s + c]; a = fma(v, v, a); } xx[tid] = a; }
    __syncthreads();

    const int qr = lane >> 2, qc = lane & 3;
    const int row0 = 16 * warp;
    double xr[2] = {xx[row0 + qr], xx[row0 + 8 + qr]};
    double ysum[KO][2], gsum[KO][2];
    double jacc[KO][2][8][2];
#pragma unroll
    for (int o = 0; o < KO; ++o)
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            ysum[o][a] = 0.0; gsum[o][a] = 0.0;
#pragma unroll
            for (int c = 0; c < 8; ++c) { jacc[o][a][c][0] = 0.0; jacc[o][a][c][1] = 0.0; }
        }
    const int ksteps = ((n + 3) & ~3) >> 2;
    const double* xa = Xs + (row0 + qr) * s + qc;

    for (int t = t_lo; t < t_hi; ++t) {
        const int it = t - t_lo, cur = it & 1;
        if (tid == 0 && t + 1 < t_hi) {
            const int nb = cur ^ 1;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar[nb])), "r"(tile_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(nb ? buf1 : buf0)), "l"(pack + (size_t)(t + 1) * tile_d), "r"(tile_bytes), "r"(smem_u32(&bar[nb])) : "memory");
        }
        {
            const unsigned parity = (unsigned)((it >> 1) & 1);
            unsigned ok = 0;
            while (!ok) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&bar[cur])), "r"(parity) : "memory");
            }
        }
        const double* Cs = cur ? buf1 : buf0;
        const double* ccs = Cs + DM_TN * s;
        const double* Wt = ccs + DM_TN + (size_t)l0 * DM_TN;
        double acc[2][8][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int c = 0; c < 8; ++c) { acc[a][c][0] = 0.0; acc[a][c][1] = 0.0; }
        const double* cb = Cs + qr * s + qc;
#pragma unroll 2
        for (int ks = 0; ks < ksteps; ++ks) {     // phase 1
            const double fa0 = xa[4 * ks], fa1 = xa[8 * s + 4 * ks];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const double fb = cb[8 * c * s + 4 * ks];
                dmma884(acc[0][c], fa0, fb);
                dmma884(acc[1][c], fa1, fb);
            }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = 8 * c + 2 * qc + e;
                const double ccv = ccs[col];
                double wv[KO];
#pragma unroll
                for (int o = 0; o < KO; ++o) wv[o] = Wt[o * DM_TN + col];
#pragma unroll
                for (int a = 0; a < 2; ++a) {
                    double r2 = fma(-2.0, acc[a][c][e], xr[a] + ccv);
                    r2 = fmax(r2, 0.0);
                    double ph, ps;
                    rad_phi_psi_t<RK>(rf, r2, ph, ps);
#pragma unroll
                    for (int o = 0; o < KO; ++o) { ysum[o][a] = fma(ph, wv[o], ysum[o][a]); gsum[o][a] = fma(ps, wv[o], gsum[o][a]); }
                    acc[a][c][e] = ps;             // psi in C-fragment layout; the output weights are applied per k-column below
                }
            }
        // phase 2: Jacc_o += (psi . w_o) (16 x 64) * C' (64 x 8 ncb)
#pragma unroll
        for (int kc = 0; kc < 16; ++kc) {
            const int c = kc >> 1, h = kc & 1;
            const int src = (lane & ~3) | (2 * h + (qc >> 1));
            double fa[2];
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                const double v0 = __shfl_sync(0xffffffffu, acc[a][c][0], src);
                const double v1 = __shfl_sync(0xffffffffu, acc[a][c][1], src);
                fa[a] = (qc & 1) ? v1 : v0;
            }
            const double* brow = Cs + (4 * kc + qc) * s + qr;
            double fo[KO][2];
#pragma unroll
            for (int o = 0; o < KO; ++o) { const double wk_ = Wt[o * DM_TN + 4 * kc + qc]; fo[o][0] = fa[0] * wk_; fo[o][1] = fa[1] * wk_; }
#pragma unroll
            for (int cbk = 0; cbk < 8; ++cbk) {
                if (cbk < ncb) {
                    const double fb = brow[8 * cbk];
#pragma unroll
                    for (int o = 0; o < KO; ++o) { dmma884(jacc[o][0][cbk], fo[o][0], fb); dmma884(jacc[o][1][cbk], fo[o][1], fb); }
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int o = 0; o < KO; ++o)
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            double v = ysum[o][a], g = gsum[o][a];
            v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2);
            g += __shfl_xor_sync(0xffffffffu, g, 1); g += __shfl_xor_sync(0xffffffffu, g, 2);
            ysum[o][a] = v; gsum[o][a] = g;
        }
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int row = row0 + 8 * a + qr;
        const long long mi = m0 + row;
        if (mi >= P.M) continue;
#pragma unroll
        for (int o = 0; o < KO; ++o) {
            const int lo_ = l0 + o;
            if (Z > 1) {
                if (P.Y && qc == 0) P.partY[(((size_t)zz * P.B + b) * P.M + mi) * k + lo_] = ysum[o][a];
                double* Jp = P.partJ + ((((size_t)zz * P.B + b) * P.M + mi) * k + lo_) * n;
#pragma unroll
                for (int cbk = 0; cbk < 8; ++cbk)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int col = 8 * cbk + 2 * qc + e;
                        if (cbk < ncb && col < n) Jp[col] = fma(Xs[row * s + col], gsum[o][a], -jacc[o][a][cbk][e]);
                    }
                continue;
            }
            if (P.Y && qc == 0) {
                double v = ysum[o][a];
                if (P.deg >= 0) v += lam[lo_];
                if (P.deg >= 1) {
                    double tsum = 0.0;
                    for (int c = 0; c < n; ++c) tsum = fma(lam[(size_t)(c + 1) * k + lo_], Xs[row * s + c] + centers[c], tsum);
                    v += tsum;
                }
                P.Y[((size_t)b * P.M + mi) * k + lo_] = v;
            }
            double* Jrow = P.J + (((size_t)b * P.M + mi) * k + lo_) * n;
#pragma unroll
            for (int cbk = 0; cbk < 8; ++cbk)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int col = 8 * cbk + 2 * qc + e;
                    if (cbk < ncb && col < n) {
                        double val = fma(Xs[row * s + col], gsum[o][a], -jacc[o][a][cbk][e]);
                        if (P.deg >= 1) val += lam[(size_t)(col + 1) * k + lo_];
                        Jrow[col] = val;
                    }
                }
        }
    }
}

template <int RK>
static cudaError_t launch_dmma_jac_t(const EvalParams& P, cudaStream_t s, int* n_launches) {
    const int st = P.pack_s;
    const size_t smem = 128 + sizeof(double) * (2 * P.pack_tile_doubles + (size_t)DM_TM * st + DM_TM);
    cudaError_t e = raise_dyn_smem(eval_dmma_jac_kernel<RK, 1>, smem);
    if (e == cudaSuccess) e = raise_dyn_smem(eval_dmma_jac_kernel<RK, 2>, smem);
    if (e != cudaSuccess) return e;
    const long long tiles = (P.M + DM_TM - 1) / DM_TM;
    for (int l0 = 0; l0 < P.k;) {                // two outputs per pass while there are two left
        dim3 grid((unsigned)tiles, (unsigned)P.B, (unsigned)(P.zsplit > 1 ? P.zsplit : 1));
        if (l0 + 1 < P.k) { eval_dmma_jac_kernel<RK, 2><<<grid, 256, smem, s>>>(P, l0); l0 += 2; }
        else { eval_dmma_jac_kernel<RK, 1><<<grid, 256, smem, s>>>(P, l0); l0 += 1; }
        if (n_launches) ++*n_launches;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}
static cudaError_t launch_dmma_jac(const EvalParams& P, cudaStream_t s, int* n_launches) {
    switch (rad_kind(P.kernel, P.ibeta)) {
    case RK_CUBIC3: return launch_dmma_jac_t<RK_CUBIC3>(P, s, n_launches);
    case RK_MQ: return launch_dmma_jac_t<RK_MQ>(P, s, n_launches);
    case RK_GAUSS: return launch_dmma_jac_t<RK_GAUSS>(P, s, n_launches);
    default: return launch_dmma_jac_t<RK_GENERIC>(P, s, n_launches);
    }
}

template <int CQ, bool WANT_J>
static cudaError_t launch_tile(const EvalParams& P, cudaStream_t s, int* n_launches) {
    constexpr int ND = 16 * CQ, KG = WANT_J ? 2 : 4;
    const size_t smem = sizeof(double) * ((size_t)ND * TM + (size_t)ND * TN + (size_t)TN * ND + (size_t)TM * GLD + TM + TN + KG * TN + ND);
    cudaError_t e = raise_dyn_smem(eval_tile_kernel<CQ, WANT_J>, smem);
    if (e != cudaSuccess) return e;
    const long long tiles = (P.M + TM - 1) / TM;
    for (int l0 = 0; l0 < P.k; l0 += KG) {
        const int kk = (P.k - l0) < KG ? (P.k - l0) : KG;
        // gridDim.x limit is 2^31-1 tiles: fine for M <= 1.3e11
        dim3 grid((unsigned)tiles, (unsigned)P.B);
        eval_tile_kernel<CQ, WANT_J><<<grid, 256, smem, s>>>(P, l0, kk);
        if (n_launches) ++*n_launches;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

static size_t eval_wide_smem(int npad) {
    return sizeof(double) * ((size_t)npad * TM + (size_t)WC * CLD + TM + TN + 4 * TN + npad);
}

static cudaError_t launch_wide(const EvalParams& P, cudaStream_t s, int* n_launches) {
    const int npad = ((P.n + WC - 1) / WC) * WC;
    const size_t smem = eval_wide_smem(npad);
    cudaError_t e = raise_dyn_smem(eval_wide_kernel, smem);
    if (e != cudaSuccess) return e;
    const long long tiles = (P.M + TM - 1) / TM;
    for (int l0 = 0; l0 < P.k; l0 += 4) {
        const int kk = (P.k - l0) < 4 ? (P.k - l0) : 4;
        dim3 grid((unsigned)tiles, (unsigned)P.B);
        eval_wide_kernel<<<grid, 256, smem, s>>>(P, l0, kk, npad);
        if (n_launches) ++*n_launches;
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}


// Sum of the zsplit partial results in a fixed order, plus the polynomial tail (values: lambda_0 + lambda' x, Jacobian: lambda).
__global__ void eval_split_reduce_kernel(EvalParams P) {
    const int n = P.n, k = P.k, pl = P.p > 0 ? P.p : 1;
    const size_t ny = (size_t)P.B * P.M * k, nj = P.J ? ny * n : 0;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < ny + nj; e += (size_t)gridDim.x * blockDim.x) {
        if (e < ny) {
            if (!P.Y) continue;
            const int l = (int)(e % k); const size_t pm = e / k; const int b = (int)(pm / P.M);
            double v = 0.0;
            for (int z = 0; z < P.zsplit; ++z) v += P.partY[(size_t)z * ny + e];
            const double* lam = P.lam + (size_t)b * pl * k;
            if (P.deg >= 0) v += lam[l];
            if (P.deg >= 1) {
                const double* x = P.X + pm * n;
                double tsum = 0.0;
                for (int c = 0; c < n; ++c) tsum = fma(lam[(size_t)(c + 1) * k + l], x[c], tsum);
                v += tsum;
            }
            P.Y[e] = v;
        } else {
            const size_t f = e - ny;
            const int col = (int)(f % n); const size_t pml = f / n; const int l = (int)(pml % k); const int b = (int)(pml / k / P.M);
            double v = 0.0;
            for (int z = 0; z < P.zsplit; ++z) v += P.partJ[(size_t)z * nj + f];
            if (P.deg >= 1) v += P.lam[(size_t)b * pl * k + (size_t)(col + 1) * k + l];
            P.J[f] = v;
        }
    }
}

cudaError_t launch_eval(const EvalParams& P, cudaStream_t s, int* n_launches) {
    if (P.M <= 0 || P.B <= 0) return cudaSuccess;
    const bool want_j = P.J != nullptr;
    if (P.deg >= 2) {                              // quadratic tail (kernels of cpd order 3): the generic kernel only
        if (P.B > 65535) return cudaErrorInvalidValue;
        dim3 grid((unsigned)((P.M + 127) / 128), (unsigned)P.B);
        eval_generic_kernel<<<grid, 128, 0, s>>>(P);
        if (n_launches) ++*n_launches;
        return cudaGetLastError();
    }
    {   // a handful of points per instance: one warp per point
        const int wd = P.n + P.train_stride * P.k;
        const long long q = (long long)P.B * P.M;
        if (P.M <= 8 && (size_t)4 * wd * sizeof(double) <= 96 * 1024 && (q + 3) / 4 < 2147483647LL) {
            const size_t smem = (size_t)4 * wd * sizeof(double);
            cudaError_t e = raise_dyn_smem(eval_small_kernel, smem);
            if (e != cudaSuccess) return e;
            eval_small_kernel<<<(unsigned)((q + 3) / 4), 128, smem, s>>>(P, wd);
            if (n_launches) ++*n_launches;
            return cudaGetLastError();
        }
    }
    if (P.pack && P.n <= 64 && P.B <= 65535 && P.k <= 16) {
        cudaError_t e = want_j ? launch_dmma_jac(P, s, n_launches) : launch_dmma(P, s, n_launches);
        if (e == cudaSuccess && P.zsplit > 1) {
            const size_t tot = (size_t)P.B * P.M * P.k * (want_j ? (size_t)(P.n + 1) : 1);
            const unsigned blocks = (unsigned)((tot + 255) / 256 < 4096 ? (tot + 255) / 256 : 4096);
            eval_split_reduce_kernel<<<blocks, 256, 0, s>>>(P);
            if (n_launches) ++*n_launches;
            e = cudaGetLastError();
        }
        return e;
    }
    if (P.n <= 64 && P.B <= 65535) {
        const int cq = (P.n + 15) / 16;
        if (want_j) {
            switch (cq) {
            case 1: return launch_tile<1, true>(P, s, n_launches);
            case 2: return launch_tile<2, true>(P, s, n_launches);
            case 3: return launch_tile<3, true>(P, s, n_launches);
            default: return launch_tile<4, true>(P, s, n_launches);
            }
        } else {
            switch (cq) {
            case 1: return launch_tile<1, false>(P, s, n_launches);
            case 2: return launch_tile<2, false>(P, s, n_launches);
            case 3: return launch_tile<3, false>(P, s, n_launches);
            default: return launch_tile<4, false>(P, s, n_launches);
            }
        }
    }
    if (P.B > 65535) return cudaErrorInvalidValue;
    if (!want_j && P.n > 64 && eval_wide_smem(((P.n + WC - 1) / WC) * WC) <= 227 * 1024) return launch_wide(P, s, n_launches);
    dim3 grid((unsigned)((P.M + 127) / 128), (unsigned)P.B);
    eval_generic_kernel<<<grid, 128, 0, s>>>(P);
    if (n_launches) ++*n_launches;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Armijo backtracking (src/descent.jl:150-185): all step sizes as one batch of trial points.
// ------------------------------------------------------------------------------------------------
__global__ void backtrack_points_kernel(BacktrackParams P) {
    const int b = blockIdx.x, n = P.n, tid = threadIdx.x, nt = blockDim.x;
    const double* x = P.x + (size_t)b * n; const double* dir = P.dir + (size_t)b * n;
    double* Xb = P.Xall + (size_t)b * (P.nsteps + 1) * n;
    double* sg = P.sig_all + (size_t)b * P.nsteps;
    for (int i = tid; i < n; i += nt) Xb[i] = x[i];
    for (int sidx = tid; sidx < P.nsteps; sidx += nt) {
        double s = P.step0[b];
        for (int i = 0; i < sidx; ++i) s *= P.shrink;        // repeated `*=` like the reference loop
        sg[sidx] = s;
    }
    __syncthreads();
    for (int e = tid; e < P.nsteps * n; e += nt) {
        const int sidx = e / n, i = e % n;
        Xb[(size_t)(sidx + 1) * n + i] = __dadd_rn(x[i], __dmul_rn(sg[sidx], dir[i]));   // x .+ step_size .* dir (no fma)
    }
}

__global__ void backtrack_pick_kernel(BacktrackParams P) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= P.B) return;
    const int n = P.n, k = P.k;
    const double* Yb = P.Yall + (size_t)b * (P.nsteps + 1) * k;
    const double* sg = P.sig_all + (size_t)b * P.nsteps;
    const double om = P.omega[b];
    int i = 0;
    // reference loop: while i < MAX_LOOPS: if armijo -> break; if step <= MIN -> break; shrink; i += 1
    for (;;) {
        bool ok;
        if (P.strict) {
            ok = true;
            for (int l = 0; l < k; ++l) ok = ok && ((Yb[l] - Yb[(size_t)(i + 1) * k + l]) >= sg[i] * P.armijo_c * om);
        } else {
            double m0 = Yb[0], m1 = Yb[(size_t)(i + 1) * k];
            for (int l = 1; l < k; ++l) { m0 = fmax(m0, Yb[l]); m1 = fmax(m1, Yb[(size_t)(i + 1) * k + l]); }
            ok = (m0 - m1) >= sg[i] * P.armijo_c * om;
        }
        if (i >= P.max_loops || ok || sg[i] <= P.min_stepsize) break;
        ++i;
    }
    P.step_index[b] = i;
    P.sigma[b] = sg[i];
    const double* Xb = P.Xall + ((size_t)b * (P.nsteps + 1) + i + 1) * n;
    for (int r = 0; r < n; ++r) P.x_plus[(size_t)b * n + r] = Xb[r];
    for (int l = 0; l < k; ++l) { P.mx[(size_t)b * k + l] = Yb[l]; P.mx_plus[(size_t)b * k + l] = Yb[(size_t)(i + 1) * k + l]; }
}

cudaError_t launch_backtrack_points(const BacktrackParams& P, cudaStream_t s) {
    backtrack_points_kernel<<<P.B, 128, 0, s>>>(P);
    return cudaGetLastError();
}
cudaError_t launch_backtrack_pick(const BacktrackParams& P, cudaStream_t s) {
    backtrack_pick_kernel<<<(P.B + 127) / 128, 128, 0, s>>>(P);
    return cudaGetLastError();
}

}  // namespace mrbf
