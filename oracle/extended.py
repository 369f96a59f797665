"""Extended-precision reference solution of the RBF saddle system (test infrastructure, like everything under oracle/).

`truth_values` solves  [Phi Pi; Pi' 0] [w; lam] = [Y; 0]  (the system RBF.RBFInterpolationModel solves with dense `\\`, assumption U5)
by iterative refinement: LU in float64, residuals and the accumulated solution in numpy.longdouble (80-bit on x86: eps 1.1e-19).  While
cond * eps64 < 1 the iteration converges to a forward error of about cond * eps_longdouble, i.e. 3 digits below anything float64 can
reach -- enough to tell WHICH of two float64 solutions (the GPU's reduced-system Cholesky, the reference's LU) is closer, and by how much.
Kernels: cubic (beta = 3) and multiquadric (beta = 1/2), polynomial degree 1."""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

LD = np.longdouble


def _phi(kernel: str, alpha: float, r2):
    if kernel == "cubic":
        return r2 * np.sqrt(r2)
    if kernel == "multiquadric":
        return -np.sqrt(LD(1) + LD(alpha) ** 2 * r2)
    raise ValueError(kernel)


def _psi(kernel: str, alpha: float, r2):       # phi'(rho) / rho
    if kernel == "cubic":
        return LD(3) * np.sqrt(r2)
    return -LD(alpha) ** 2 / np.sqrt(LD(1) + LD(alpha) ** 2 * r2)


def truth_values(kernel: str, alpha: float, S, V, X, iters: int = 12):
    """Returns (Y, J, info): model values (M x k) and Jacobians (M x k x n) at X in longdouble, info = dict(cond, residual, steps)."""
    S = np.asarray(S, np.float64); V = np.asarray(V, np.float64); X = np.asarray(X, np.float64)
    N, n = S.shape; k = V.shape[1]
    Sl = S.astype(LD)
    d = Sl[:, None, :] - Sl[None, :, :]
    Phi = _phi(kernel, alpha, np.sum(d * d, -1))
    Pi = np.concatenate([np.ones((N, 1), LD), Sl], 1)
    K = np.zeros((N + n + 1, N + n + 1), LD)
    K[:N, :N] = Phi; K[:N, N:] = Pi; K[N:, :N] = Pi.T
    rhs = np.zeros((N + n + 1, k), LD); rhs[:N] = V
    K64 = K.astype(np.float64)
    lu = sla.lu_factor(K64)
    x = sla.lu_solve(lu, rhs.astype(np.float64)).astype(LD)
    res = np.inf; steps = 0
    for steps in range(1, iters + 1):
        r = rhs - K @ x
        res_new = float(np.max(np.abs(r)))
        if not res_new < 0.5 * res:          # stagnation: converged to the longdouble residual level
            break
        res = res_new
        x = x + sla.lu_solve(lu, r.astype(np.float64)).astype(LD)
    w, lam = x[:N], x[N:]
    Xl = X.astype(LD)
    dx = Xl[:, None, :] - Sl[None, :, :]
    r2 = np.sum(dx * dx, -1)
    Y = _phi(kernel, alpha, r2) @ w + np.concatenate([np.ones((len(X), 1), LD), Xl], 1) @ lam
    ps = _psi(kernel, alpha, r2)                                   # M x N
    J = np.einsum("mn,mnc,nk->mkc", ps, dx, w) + lam[1:].T[None, :, :]
    return Y, J, dict(cond=float(np.linalg.cond(K64)), residual=res, steps=steps)
