"""TEST INFRASTRUCTURE ONLY (CPU oracle; never imported by the product path).

Restatement of the linear programme behind Morbit's constrained steepest-descent direction
(`_steepest_descent_direction`, /root/reference/src/descent.jl:75-135):

        min_{alpha, d}  alpha   s.t.   Df_i . d <= alpha * ||Df_i||_2   (rows normalised iff `normalize`, :112-116)
                                       -1 <= d <= 1                      (:117)
                                       lb <= x + d <= ub                 (:120)
        returns (d, omega = -alpha)                                      (:134)

The reference hands this to JuMP with OSQP (`eps_rel = 1e-5`, `polish = true`, :101-102), a dependency that is not vendored
(Project.toml) and whose iterates are only 1e-5-accurate, so there are no golden numbers to pin: parity unpinned.  The LP itself
is pinned mathematically, and two independent exact solvers restate it here:
  * `lp_highs`   -- SciPy's HiGHS (`scipy.optimize.linprog`), general k;
  * `lp_k2_exact`-- for k = 2 the dual is a concave piecewise-linear function of one multiplier; enumerate its breakpoints.
"""
from __future__ import annotations

import numpy as np


def _rows(jac, normalize):
    jac = np.asarray(jac, dtype=np.float64)
    nrm = np.linalg.norm(jac, axis=1) if normalize else np.ones(jac.shape[0])
    return jac, nrm


def box(x, lb, ub):
    x = np.asarray(x, dtype=np.float64)
    return np.maximum(-1.0, np.asarray(lb) - x), np.minimum(1.0, np.asarray(ub) - x)


def lp_highs(x, jac, lb, ub, normalize=True):
    """(d, omega) from HiGHS."""
    from scipy.optimize import linprog
    jac, nrm = _rows(jac, normalize)
    k, n = jac.shape
    lo, hi = box(x, lb, ub)
    if not np.any(nrm > 0):
        return np.zeros(n), 0.0
    c = np.zeros(n + 1); c[n] = 1.0
    A = np.hstack([jac, -nrm[:, None]])
    bounds = [(lo[j], hi[j]) for j in range(n)] + [(None, None)]
    res = linprog(c, A_ub=A, b_ub=np.zeros(k), bounds=bounds, method="highs")
    assert res.status == 0, res.message
    return res.x[:n], -float(res.x[n])


def dual_value(lmbda, jac, nrm, lo, hi):
    """phi(lambda) = min over the box of (sum_i lambda_i g_i) . d, g_i = Df_i / nrm_i."""
    g = (jac / np.where(nrm > 0, nrm, 1.0)[:, None])
    c = lmbda @ g
    return float(np.sum(np.minimum(c * lo, c * hi)))


def lp_k2_exact(x, jac, lb, ub, normalize=True):
    """omega for k = 2 by enumerating the breakpoints of the concave piecewise-linear dual (alpha* = max_lambda phi)."""
    jac, nrm = _rows(jac, normalize)
    assert jac.shape[0] == 2 and np.all(nrm > 0)
    lo, hi = box(x, lb, ub)
    g1, g2 = jac[0] / nrm[0], jac[1] / nrm[1]
    cands = [0.0, 1.0]
    den = g1 - g2
    with np.errstate(divide="ignore", invalid="ignore"):
        t = -g2 / den                      # t g1 + (1 - t) g2 = 0
    cands += [float(v) for v in t[np.isfinite(t) & (t > 0) & (t < 1)]]
    best = max(dual_value(np.array([t_, 1.0 - t_]), jac, nrm, lo, hi) for t_ in cands)
    return -best


def check_optimal(x, jac, lb, ub, d, omega, normalize=True, tol=1e-9):
    """Primal feasibility of (d, alpha = -omega) and optimality against HiGHS."""
    jac, nrm = _rows(jac, normalize)
    lo, hi = box(x, lb, ub)
    assert np.all(d >= lo - tol) and np.all(d <= hi + tol), "d leaves the box"
    if np.any(nrm > 0):
        assert np.all(jac @ d <= -omega * nrm + tol * (1.0 + np.abs(jac).sum(axis=1))), "descent rows violated"
    _, om_ref = lp_highs(x, jac, lb, ub, normalize)
    assert abs(omega - om_ref) <= tol * max(1.0, abs(om_ref)), (omega, om_ref)
