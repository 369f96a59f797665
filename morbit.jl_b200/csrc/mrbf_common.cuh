// Shared device helpers for the sm_100a RBF kernels.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/morbit_rbf.h"

namespace mrbf {

// Radial function with resolved parameters (what RBF._get_rad_func returns, RbfModel.jl:692-696).
struct RadFn {
    int kernel;      // enum mrbf_kernel
    int ibeta;       // cubic exponent (1,3,5,..) or thin-plate k
    double alpha2;   // shape parameter squared
    double sgn;      // sign making phi conditionally positive definite
};

// phi as a function of the SQUARED distance (saves the sqrt for gaussian / multiquadrics).
__device__ __forceinline__ double rad_phi(const RadFn& f, double r2) {
    switch (f.kernel) {
    case MRBF_CUBIC: {
        double r = sqrt(r2);
        double out = r;
        for (int e = 2; e < f.ibeta; e += 2) out *= r2;      // r^beta for odd beta
        return f.sgn * out;
    }
    case MRBF_MULTIQUADRIC: return f.sgn * sqrt(fma(f.alpha2, r2, 1.0));
    case MRBF_INV_MULTIQUADRIC: return 1.0 / sqrt(fma(f.alpha2, r2, 1.0));
    case MRBF_GAUSSIAN: return exp(-f.alpha2 * r2);
    default: {   // thin plate spline: sgn * r^(2k) log r
        if (r2 == 0.0) return 0.0;
        double p = r2;
        for (int e = 1; e < f.ibeta; ++e) p *= r2;
        return f.sgn * p * (0.5 * log(r2));
    }
    }
}

// psi = phi'(r)/r as a function of the squared distance; 0 where singular at r = 0.
__device__ __forceinline__ double rad_psi(const RadFn& f, double r2) {
    switch (f.kernel) {
    case MRBF_CUBIC: {
        if (r2 == 0.0) return 0.0;
        double r = sqrt(r2);
        double out = (f.ibeta == 1) ? 1.0 / r : r;            // beta r^(beta-2)
        for (int e = 4; e < f.ibeta; e += 2) out *= r2;
        return f.sgn * (double)f.ibeta * out;
    }
    case MRBF_MULTIQUADRIC: return f.sgn * f.alpha2 / sqrt(fma(f.alpha2, r2, 1.0));
    case MRBF_INV_MULTIQUADRIC: {
        double t = fma(f.alpha2, r2, 1.0);
        return -f.alpha2 / (t * sqrt(t));
    }
    case MRBF_GAUSSIAN: return -2.0 * f.alpha2 * exp(-f.alpha2 * r2);
    default: {
        if (r2 == 0.0) return 0.0;
        double p = 1.0;
        for (int e = 1; e < f.ibeta; ++e) p *= r2;
        return f.sgn * p * ((double)f.ibeta * log(r2) + 1.0);
    }
    }
}

// phi and psi together (shares the transcendental).
__device__ __forceinline__ void rad_phi_psi(const RadFn& f, double r2, double& phi, double& psi) {
    switch (f.kernel) {
    case MRBF_CUBIC: {
        double r = sqrt(r2);
        if (f.ibeta == 3) { phi = f.sgn * r2 * r; psi = f.sgn * 3.0 * r; return; }
        phi = rad_phi(f, r2); psi = rad_psi(f, r2); return;
    }
    case MRBF_MULTIQUADRIC: {
        double s = sqrt(fma(f.alpha2, r2, 1.0));
        phi = f.sgn * s; psi = f.sgn * f.alpha2 / s; return;
    }
    case MRBF_INV_MULTIQUADRIC: {
        double t = fma(f.alpha2, r2, 1.0);
        double is = 1.0 / sqrt(t);
        phi = is; psi = -f.alpha2 * is / t; return;
    }
    case MRBF_GAUSSIAN: {
        double e = exp(-f.alpha2 * r2);
        phi = e; psi = -2.0 * f.alpha2 * e; return;
    }
    default: phi = rad_phi(f, r2); psi = rad_psi(f, r2); return;
    }
}

// Reciprocal and reciprocal square root: hardware seed (about 20 bits) + Newton steps, accurate to a few ulp.  For normal positive
// arguments; the full IEEE division / sqrt sequences cost several times as many FP64-pipe instructions.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0); r = fma(r, e, r);
    e = fma(-x, r, 1.0); r = fma(r, e, r);
    return r;
}
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double h = 0.5 * x;
    double e = fma(-h * y, y, 0.5); y = fma(y, e, y);
    e = fma(-h * y, y, 0.5); y = fma(y, e, y);
    e = fma(-h * y, y, 0.5); y = fma(y, e, y);
    return y;
}

// Radial functions with the kernel fixed at COMPILE time for the throughput sweeps (no switch, no exponent loop in the epilogue
// of every point-centre pair).  RK_GENERIC falls back to the run-time versions.
enum { RK_GENERIC = 0, RK_CUBIC3 = 1, RK_MQ = 2, RK_GAUSS = 3 };
__host__ __device__ inline int rad_kind(int kernel, int ibeta) {
    if (kernel == MRBF_CUBIC && ibeta == 3) return RK_CUBIC3;
    if (kernel == MRBF_MULTIQUADRIC) return RK_MQ;
    if (kernel == MRBF_GAUSSIAN) return RK_GAUSS;
    return RK_GENERIC;
}
template <int RK>
__device__ __forceinline__ double rad_phi_t(const RadFn& f, double r2) {
    if constexpr (RK == RK_CUBIC3) {             // rho^3 = r2 * sqrt(r2)   (sign (-1)^2 = +1)
        const double r = (r2 > 0.0) ? r2 * fast_rsqrt(r2) : 0.0;
        return r2 * r;
    } else if constexpr (RK == RK_MQ) {          // -sqrt(1 + (alpha rho)^2), argument >= 1
        const double t = fma(f.alpha2, r2, 1.0);
        return -(t * fast_rsqrt(t));
    } else if constexpr (RK == RK_GAUSS) {
        return exp(-f.alpha2 * r2);
    } else {
        return rad_phi(f, r2);
    }
}
template <int RK>
__device__ __forceinline__ void rad_phi_psi_t(const RadFn& f, double r2, double& phi, double& psi) {
    if constexpr (RK == RK_CUBIC3) {
        const double r = (r2 > 0.0) ? r2 * fast_rsqrt(r2) : 0.0;
        phi = r2 * r; psi = 3.0 * r;
    } else if constexpr (RK == RK_MQ) {
        const double t = fma(f.alpha2, r2, 1.0);
        const double rs = fast_rsqrt(t);
        phi = -(t * rs); psi = -f.alpha2 * rs;
    } else if constexpr (RK == RK_GAUSS) {
        const double e = exp(-f.alpha2 * r2);
        phi = e; psi = -2.0 * f.alpha2 * e;
    } else {
        rad_phi_psi(f, r2, phi, psi);
    }
}

// results_in_box_indices' test (Databases.jl:324-327): inclusive bounds on every coordinate.
__device__ __forceinline__ bool in_box_pt(const double* s, const double* lb, const double* ub, int n) {
    bool ok = true;
    for (int i = 0; i < n; ++i) ok = ok && (lb[i] <= s[i]) && (s[i] <= ub[i]);
    return ok;
}

// intersect_box(x, d, lb, ub; return_vals = :absmax), src/utilities.jl:126-221.
__device__ inline double intersect_box_absmax(int n, const double* x, const double* d, const double* lb, const double* ub) {
    bool any = false;
    for (int i = 0; i < n; ++i) any = any || (d[i] != 0.0);
    if (!any) return INFINITY;
    double s_pos = 0.0, s_neg = 0.0;
    bool have_pos = false, have_neg = false;
    for (int pass = 0; pass < 2; ++pass)
        for (int i = 0; i < n; ++i) {
            if (d[i] == 0.0) continue;
            double tmp = (pass == 0 ? lb[i] : ub[i]) - x[i];
            double sig;
            if (tmp != 0.0) sig = tmp / d[i];
            else if (pass == 0) sig = d[i] > 0.0 ? INFINITY : 0.0;
            else sig = d[i] < 0.0 ? INFINITY : 0.0;
            if (sig >= 0.0) { if (!have_pos || sig < s_pos) s_pos = sig; have_pos = true; }
            else { if (!have_neg || sig > s_neg) s_neg = sig; have_neg = true; }
        }
    if (!have_pos) s_pos = 0.0;
    if (!have_neg) s_neg = 0.0;
    return fabs(s_pos) >= fabs(s_neg) ? s_pos : s_neg;
}

// dimension of the polynomial tail: none, constants, linear, or full quadratic (monomials 1, x_1..x_n, x_i x_j with i <= j, i outer).
// Degree 2 only arises when a kernel's order of conditional positive definiteness raises the configured degree (thin plate spline
// k = 2, cubic beta = 5): the build and the generic evaluation kernel handle it, round 4 always works with the configured degree.
__host__ __device__ inline int poly_dim(int n, int deg) { return deg < 0 ? 0 : (deg == 0 ? 1 : (deg == 1 ? n + 1 : ((n + 1) * (n + 2)) / 2)); }
// value of basis function c (0-based) of the degree-<= 2 monomial basis at x
__device__ __forceinline__ double poly_basis_at(const double* x, int n, int c) {
    if (c == 0) return 1.0;
    if (c <= n) return x[c - 1];
    int a = 0, rem = c - (n + 1);
    while (rem >= n - a) { rem -= n - a; ++a; }
    return x[a] * x[a + rem];
}

// ---- block-wide reductions (deterministic order) ----------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Sum over the block; result returned to every thread.  red: >= 33 doubles of shared scratch.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();                       // protect red[] from the previous use
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double t = (lane < nw) ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}

// Two sums at once.
__device__ __forceinline__ void block_sum2(double& a, double& b, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    a = warp_sum(a); b = warp_sum(b);
    __syncthreads();
    if (lane == 0) { red[warp] = a; red[33 + warp] = b; }
    __syncthreads();
    if (warp == 0) {
        double t = (lane < nw) ? red[lane] : 0.0;
        double u = (lane < nw) ? red[33 + lane] : 0.0;
        t = warp_sum(t); u = warp_sum(u);
        if (lane == 0) { red[32] = t; red[65] = u; }
    }
    __syncthreads();
    a = red[32]; b = red[65];
}

// argmax with "first maximiser" tie-breaking (smallest id wins), ids >= 0; id = -1 when empty.
struct ArgMax { double v; int id; };
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {
    if (b.id < 0) return a;
    if (a.id < 0) return b;
    if (b.v > a.v || (b.v == a.v && b.id < a.id)) return b;
    return a;
}
__device__ __forceinline__ ArgMax block_argmax(ArgMax m, double* redv, int* redi) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgMax t; t.v = __shfl_xor_sync(0xffffffffu, m.v, o); t.id = __shfl_xor_sync(0xffffffffu, m.id, o);
        m = better(m, t);
    }
    __syncthreads();
    if (lane == 0) { redv[warp] = m.v; redi[warp] = m.id; }
    __syncthreads();
    if (warp == 0) {
        ArgMax t; t.v = (lane < nw) ? redv[lane] : 0.0; t.id = (lane < nw) ? redi[lane] : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ArgMax s; s.v = __shfl_xor_sync(0xffffffffu, t.v, o); s.id = __shfl_xor_sync(0xffffffffu, t.id, o);
            t = better(t, s);
        }
        if (lane == 0) { redv[32] = t.v; redi[32] = t.id; }
    }
    __syncthreads();
    ArgMax r; r.v = redv[32]; r.id = redi[32];
    return r;
}

// LinearAlgebra.givensAlgorithm convention (utilities.jl:443 -> LAPACK dlartg, pre-3.10 signs).
__device__ __forceinline__ void givens(double f, double g, double& c, double& s) {
    if (g == 0.0) { c = 1.0; s = 0.0; return; }
    if (f == 0.0) { c = 0.0; s = 1.0; return; }
    double r = hypot(f, g);
    c = f / r; s = g / r;
    if (fabs(f) > fabs(g) && c < 0.0) { c = -c; s = -s; }
}

}  // namespace mrbf
