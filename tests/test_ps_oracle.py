"""CPU: the restated Pascoletti-Serafini / ideal-point solver (oracle/ps_oracle.py) against an independent multi-start SQP solve of
the smooth problem the reference hands to NLopt (src/descent.jl:435-500).  NLopt's ISRES stream cannot be reproduced (parity unpinned
for this row), so the bar is solution quality: the evolution strategy with the reference's budget (500 (n + 1) evaluations) must land
within a stated distance of the optimum and never return an infeasible point."""
import numpy as np
import pytest

from oracle import ps_oracle as PS
from oracle import rbf_oracle as O


def _model(n, k, kernel, seed, N=None):
    rng = np.random.default_rng(seed)
    N = N or 3 * n + 4
    S = rng.random((N, n))
    a = rng.random((k, n))
    V = np.stack([np.sum((S - a[l]) ** 2, -1) for l in range(k)], -1)
    return O.build_model(S, V, O.RbfConfig(kernel=kernel)), S


def _evalf(m):
    return lambda X: np.array([m.eval(x) for x in np.atleast_2d(X)])


def test_hash_random_numbers_are_uniform_and_normal():
    u = np.array([PS.u01(PS.key(3, 1, 2, i, 0, 0)) for i in range(4000)])
    z = np.array([PS.normal01(PS.key(3, 1, 2, i, 1, 3)) for i in range(4000)])
    assert 0.0 < u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.02 and abs(u.var() - 1 / 12) < 0.01
    assert abs(z.mean()) < 0.06 and abs(z.std() - 1.0) < 0.05


@pytest.mark.parametrize("n,kernel", [(2, "cubic"), (3, "multiquadric")])
def test_ps_solution_quality_vs_sqp(n, kernel):
    m, S = _model(n, 2, kernel, 11 * n)
    rng = np.random.default_rng(n)
    for trial in range(3):
        x = 0.2 + 0.6 * rng.random(n)
        delta = 0.15
        lb, ub = np.maximum(0.0, x - delta), np.minimum(1.0, x + delta)
        mx = m.eval(x)
        ideal = np.array([PS.ps_solve(_evalf(m), x, lb, ub, None, None, 2, l, seed=5 + l)[0] for l in range(2)])
        ideal_ref = np.array([PS.reference_optimum(_evalf(m), m.jac, x, lb, ub, None, None, 2, l) for l in range(2)])
        scale = np.abs(mx - ideal_ref).max()
        assert np.all(ideal >= ideal_ref - 1e-9 * max(1.0, scale))          # never below the true minimum
        assert np.all(ideal - ideal_ref <= 2e-2 * scale + 1e-12)             # xtol_rel = 1e-3 class accuracy of the reference
        r = mx - ideal_ref
        if np.any(r <= 0):
            continue
        tau, xm, ym, found, evals = PS.ps_solve(_evalf(m), x, lb, ub, mx, r, 2, -1, seed=9)
        tau_ref = PS.reference_optimum(_evalf(m), m.jac, x, lb, ub, mx, r, 2)
        assert found == 1 and evals <= 500 * (n + 1) + 20 * (n + 1)
        assert -1.0 <= tau <= 0.0 and tau >= tau_ref - 1e-9
        assert tau - tau_ref <= 3e-2 * abs(tau_ref) + 1e-9, (tau, tau_ref)
        # the returned point is feasible for the reference's constraints (descent.jl:435-448) and inside the box
        assert np.all(m.eval(xm) - mx - tau * r <= 1e-12) and np.all(xm >= lb) and np.all(xm <= ub)
        np.testing.assert_allclose(ym, m.eval(xm), rtol=0, atol=1e-12)


def test_constraint_surrogate_is_respected():
    """An extra output acts as c(xi) <= 0 (get_nl_ineq_constraints_optim_handles): the unconstrained minimiser is cut off."""
    n = 2
    rng = np.random.default_rng(0)
    S = rng.random((14, n))
    V = np.stack([np.sum((S - 0.9) ** 2, -1), np.sum((S - np.array([0.9, 0.1])) ** 2, -1), S[:, 0] + S[:, 1] - 1.0], -1)   # c = x1 + x2 - 1
    m = O.build_model(S, V, O.RbfConfig(kernel="cubic"))
    x = np.array([0.4, 0.4]); lb, ub = x - 0.3, x + 0.3
    f, xm, ym, found, _ = PS.ps_solve(_evalf(m), x, lb, ub, None, None, 2, 0, seed=1)
    ref = PS.reference_optimum(_evalf(m), m.jac, x, lb, ub, None, None, 2, 0)
    assert found == 1 and ym[2] <= 0.0 and xm.sum() <= 1.0 + 1e-9
    assert f >= ref - 1e-9 and f - ref <= 2e-2 * abs(m.eval(x)[0] - ref)


def test_critical_point_returns_the_iterate():
    """At a point where no xi improves every objective, tau = 0 and the start point itself is the best individual."""
    n = 2
    g = np.linspace(0.0, 1.0, 5)
    S = np.array([[a, b] for a in g for b in g])
    V = np.stack([np.sum((S - 0.0) ** 2, -1), np.sum((S - 1.0) ** 2, -1)], -1)
    m = O.build_model(S, V, O.RbfConfig(kernel="cubic"))
    x = np.array([0.5, 0.5]); lb, ub = x - 0.1, x + 0.1          # Pareto critical for the two paraboloids
    mx = m.eval(x)
    tau, xm, ym, found, _ = PS.ps_solve(_evalf(m), x, lb, ub, mx, np.array([1.0, 1.0]), 2, -1, seed=2, max_evals=600)
    assert found == 1 and abs(tau) <= 1e-3
