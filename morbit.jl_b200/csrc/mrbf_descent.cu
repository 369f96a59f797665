// Constrained steepest-descent direction of the surrogate Jacobian (src/descent.jl:75-135), batched: one warp per instance.
//
// The reference builds, per iteration and per optimize() call, a JuMP model and hands it to OSQP (eps_rel = 1e-5):
//        min alpha   s.t.   Df_i . d <= alpha * ||Df_i||  (rows normalised when `normalize`),   -1 <= d <= 1,   lb <= x + d <= ub
// and returns (d, omega = -alpha).  This is a linear programme in n + 1 variables with k general rows and a box, so it is
// solved here EXACTLY by a bounded-variable primal simplex whose basis is only k x k (k = number of surrogate outputs):
//   variables  d_1..d_n (boxed), alpha (free, always basic), slacks s_i = alpha * nrm_i - Df_i . d >= 0;   rows  Df d - alpha nrm + s = 0
//   start      d at the box vertex that minimises the mean normalised gradient, alpha = max_i Df_i . d / nrm_i, basis {alpha, s_i (i != argmax)}
//   iteration  duals y = row of B^-1 that belongs to alpha; reduced costs of the non-basic d_j: -y . Df_j (lanes over j), of the slacks: -y_i;
//              Dantzig entering rule (largest violation, smallest index on ties), ratio test with bound flips, explicit B^-1 update.
// The optimum value (hence omega) is unique; where the optimal face is degenerate the vertex returned may differ from OSQP's
// interior-ish point by more than its 1e-5 tolerance -- callers that need parity compare omega and the optimality of d.
// linear equality / inequality constraints of the MOP (A_eq, A_ineq) are outside the hot path (MRBF configs have none).
#include "mrbf_common.cuh"
#include "mrbf_kernels.h"

namespace mrbf {

constexpr int DK = 8;              // max number of outputs (rows of Df)

__global__ void __launch_bounds__(128) descent_direction_kernel(DescentParams P) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int b = blockIdx.x * (blockDim.x >> 5) + wib;
    if (b >= P.B) return;
    const int n = P.n, k = P.k;
    // per-warp shared state
    double* ws = smem + (size_t)wib * P.warp_doubles;
    double* dv = ws;                     // n      current d
    double* lo = dv + n;                 // n
    double* hi = lo + n;                 // n
    double* Binv = hi + n;               // k x k  (row-major)
    double* xB = Binv + DK * DK;         // k      values of the basic variables
    double* wcol = xB + DK;              // k      B^-1 a_q
    double* nrm = wcol + DK;             // k
    int* basis = reinterpret_cast<int*>(nrm + DK);      // k: 0..n-1 = d_j, n = alpha, n+1+i = s_i
    int* where_s = basis + DK;                           // k: position of slack i in the basis or -1 (non-basic at 0)
    unsigned char* dstat = reinterpret_cast<unsigned char*>(where_s + DK);   // n: 0 at lower, 1 at upper, 2 basic
    const double* G = P.jac + (size_t)b * k * n;         // Df, k x n row-major
    const double* x = P.x + (size_t)b * n;
    double* d_out = P.d + (size_t)b * n;

    // box: max(-1, lb - x) <= d <= min(1, ub - x)
    for (int j = lane; j < n; j += 32) {
        lo[j] = fmax(-1.0, P.lb[j] - x[j]);
        hi[j] = fmin(1.0, P.ub[j] - x[j]);
    }
    for (int i = lane; i < k; i += 32) {
        double s = 0.0;
        for (int j = 0; j < n; ++j) s = fma(G[i * n + j], G[i * n + j], s);
        nrm[i] = P.normalize ? sqrt(s) : 1.0;
    }
    __syncwarp();
    int nact = 0;                                        // rows with a non-zero right-hand side scale
    for (int i = 0; i < k; ++i) nact += (nrm[i] > 0.0) ? 1 : 0;
    if (nact == 0) {                                     // every gradient vanishes: critical point
        for (int j = lane; j < n; j += 32) d_out[j] = 0.0;
        if (lane == 0) { P.omega[b] = 0.0; if (P.iters) P.iters[b] = 0; if (P.status) P.status[b] = 0; }
        return;
    }
    // start vertex: minimise the mean normalised gradient over the box
    for (int j = lane; j < n; j += 32) {
        double c = 0.0;
        for (int i = 0; i < k; ++i) if (nrm[i] > 0.0) c += G[i * n + j] / nrm[i];
        const bool up = c < 0.0;
        dv[j] = up ? hi[j] : lo[j];
        dstat[j] = up ? 1 : 0;
    }
    __syncwarp();
    // alpha = max_i Df_i . d / nrm_i over the active rows; basis = {alpha} + slacks of the other rows
    double alpha = -INFINITY; int i0 = -1;
    for (int i = 0; i < k; ++i) {
        double part = 0.0;
        for (int j = lane; j < n; j += 32) part = fma(G[i * n + j], dv[j], part);
        part = warp_sum(part);
        if (lane == 0) wcol[i] = part;                   // Df_i . d
        if (nrm[i] > 0.0) { const double a = part / nrm[i]; if (a > alpha) { alpha = a; i0 = i; } }
    }
    __syncwarp();
    if (lane == 0) {
        // basis position r holds: r == i0 -> alpha, else slack r.  B columns: alpha -> -nrm, slack r -> e_r.
        // B^-1: rows/cols of the slacks are identity; column i0 of B is -nrm, so row i0 of B^-1 is -e_i0 / nrm_i0 and
        // the other rows get  B^-1[r][i0] = -nrm_r / nrm_i0.
        for (int r = 0; r < k; ++r)
            for (int c = 0; c < k; ++c) Binv[r * DK + c] = (r == c && r != i0) ? 1.0 : 0.0;
        for (int r = 0; r < k; ++r) Binv[r * DK + i0] = (r == i0) ? -1.0 / nrm[i0] : -nrm[r] / nrm[i0];
        for (int r = 0; r < k; ++r) {
            basis[r] = (r == i0) ? n : n + 1 + r;
            where_s[r] = (r == i0) ? -1 : r;
            xB[r] = (r == i0) ? alpha : alpha * nrm[r] - wcol[r];
        }
    }
    __syncwarp();
    const double tol = 1e-11;
    int it = 0, status = 0;
    const int max_it = 50 * (n + k) + 100;
    for (; it < max_it; ++it) {
        // duals: y = row of B^-1 at alpha's basis position
        int pa = 0;
        for (int r = 0; r < k; ++r) if (basis[r] == n) pa = r;
        // entering variable: largest violation of the reduced-cost sign condition
        double bestv = tol; int beste = -1;              // beste: 0..n-1 d_j, n+1+i slack i
        for (int j = lane; j < n; j += 32) {
            const unsigned st = dstat[j];
            if (st == 2 || !(hi[j] > lo[j])) continue;
            double r = 0.0;
            for (int i = 0; i < k; ++i) r = fma(-Binv[pa * DK + i], G[i * n + j], r);
            const double viol = (st == 0) ? -r : r;      // at lower: want r >= 0; at upper: want r <= 0
            if (viol > bestv) { bestv = viol; beste = j; }
        }
        if (lane < k && where_s[lane] < 0) {             // non-basic slack (at 0): reduced cost -y_i must be >= 0
            const double viol = Binv[pa * DK + lane];
            if (viol > bestv) { bestv = viol; beste = n + 1 + lane; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bestv, o); const int oe = __shfl_xor_sync(0xffffffffu, beste, o);
            if (oe >= 0 && (beste < 0 || ov > bestv || (ov == bestv && oe < beste))) { bestv = ov; beste = oe; }
        }
        if (beste < 0) break;                            // optimal
        const bool is_d = beste < n;
        const int si = is_d ? -1 : beste - n - 1;
        const double sigma = (is_d && dstat[beste] == 1) ? -1.0 : 1.0;       // direction of the entering variable
        // w = B^-1 a_q
        if (lane < k) {
            double a = 0.0;
            if (is_d) for (int i = 0; i < k; ++i) a = fma(Binv[lane * DK + i], G[i * n + beste], a);
            else a = Binv[lane * DK + si];
            wcol[lane] = a;
        }
        __syncwarp();
        // ratio test (lane 0; k is tiny): basic x_r moves by -sigma * t * w_r
        double tmax = is_d ? hi[beste] - lo[beste] : INFINITY; int leave = -1; int leave_at_upper = 0;
        for (int r = 0; r < k; ++r) {
            const int v = basis[r];
            if (v == n) continue;                        // alpha is free
            const double dw = sigma * wcol[r];
            double lim = INFINITY; int atu = 0;
            if (v > n) { if (dw > 1e-13) lim = fmax(xB[r], 0.0) / dw; }                            // slack >= 0
            else {
                if (dw > 1e-13) lim = fmax(xB[r] - lo[v], 0.0) / dw;
                else if (dw < -1e-13) { lim = fmax(hi[v] - xB[r], 0.0) / (-dw); atu = 1; }
            }
            if (lim < tmax) { tmax = lim; leave = r; leave_at_upper = atu; }
        }
        if (!(tmax < INFINITY)) { status = 2; break; }   // unbounded: cannot happen for a bounded box
        // move
        if (lane < k) xB[lane] -= sigma * tmax * wcol[lane];
        __syncwarp();
        if (leave < 0) {                                 // bound flip of the entering d_j
            if (lane == 0) { dv[beste] = (sigma > 0.0) ? hi[beste] : lo[beste]; dstat[beste] = (sigma > 0.0) ? 1 : 0; }
        } else {
            const int lv = basis[leave];
            const double enter_val = (is_d ? dv[beste] : 0.0) + sigma * tmax;
            const double piv = wcol[leave];
            // B^-1 update: row `leave` /= piv, the others -= w_r * that row
            if (lane < k) {
                const double prow = Binv[leave * DK + lane] / piv;
                for (int r = 0; r < k; ++r) if (r != leave) Binv[r * DK + lane] = fma(-wcol[r], prow, Binv[r * DK + lane]);
                Binv[leave * DK + lane] = prow;
            }
            __syncwarp();
            if (lane == 0) {
                if (lv > n) where_s[lv - n - 1] = -1;
                else { dv[lv] = leave_at_upper ? hi[lv] : lo[lv]; dstat[lv] = leave_at_upper ? 1 : 0; }
                basis[leave] = beste; xB[leave] = enter_val;
                if (is_d) dstat[beste] = 2; else where_s[si] = leave;
            }
        }
        __syncwarp();
    }
    if (it >= max_it) status = 1;
    // write back: basic d_j take their basis value
    if (lane == 0) for (int r = 0; r < k; ++r) if (basis[r] < n) dv[basis[r]] = xB[r];
    __syncwarp();
    for (int j = lane; j < n; j += 32) d_out[j] = dv[j];
    if (lane == 0) {
        double a = 0.0;
        for (int r = 0; r < k; ++r) if (basis[r] == n) a = xB[r];
        P.omega[b] = -a;
        if (P.iters) P.iters[b] = it;
        if (P.status) P.status[b] = status;
    }
}

size_t descent_warp_doubles(int n, int k) {
    (void)k;
    size_t d = 3 * (size_t)n + DK * DK + 3 * DK + DK /* basis + where_s ints */ + ((size_t)n + 7) / 8 + 2;
    return (d + 1) & ~(size_t)1;
}
int descent_max_outputs() { return DK; }

cudaError_t launch_descent_direction(const DescentParams& P, cudaStream_t s) {
    const int wpb = 4;
    const size_t smem = wpb * P.warp_doubles * sizeof(double);
    cudaError_t e = raise_dyn_smem(descent_direction_kernel, smem);
    if (e != cudaSuccess) return e;
    descent_direction_kernel<<<(P.B + wpb - 1) / wpb, wpb * 32, smem, s>>>(P);
    return cudaGetLastError();
}

}  // namespace mrbf
