"""GPU: mrbf_descent_direction (the LP of src/descent.jl:75-135, solved exactly) against the CPU oracle (HiGHS / breakpoint
enumeration).  omega is unique and must agree to 1e-9; d must be primal feasible and optimal (the optimal face may be degenerate,
so d itself is only compared where the LP has a unique vertex solution)."""
import numpy as np
import pytest

import morbit_jl_b200 as mb
from oracle import descent_oracle as D

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    return mb.Engine(0)


@pytest.mark.parametrize("n,k,boxed,normalize", [(2, 2, False, True), (30, 2, True, True), (30, 2, True, False), (10, 1, True, True),
                                                  (200, 5, True, True), (7, 8, True, True), (30, 3, False, True)])
def test_descent_direction_matches_lp_oracle(engine, n, k, boxed, normalize):
    rng = np.random.default_rng(n * 100 + k)
    B = 33
    jac = rng.normal(size=(B, k, n)) * rng.choice([1.0, 1e-4, 30.0], size=(B, 1, 1))
    x = rng.random((B, n))
    lb, ub = (np.zeros(n), np.ones(n)) if boxed else (np.full(n, -np.inf), np.full(n, np.inf))
    if boxed:
        x[::4, 0] = 0.0; x[1::4, -1] = 1.0           # iterates on the boundary
        jac[::5, :, 1] = 0.0                          # a coordinate no output depends on (degenerate column)
    jac[2, 0, :] = 0.0                                # one output with a vanishing gradient
    if k > 1:
        jac[3, 1] = -2.0 * jac[3, 0]                  # conflicting objectives: omega = 0 (Pareto critical)
    d, omega, iters, status = engine.descent_direction(jac, x, lb, ub, normalize)
    assert np.all(status == 0), status
    for b in range(B):
        D.check_optimal(x[b], jac[b], lb, ub, d[b], omega[b], normalize, tol=1e-9)
        if k == 2 and np.all(np.linalg.norm(jac[b], axis=1) > 0):
            assert abs(omega[b] - D.lp_k2_exact(x[b], jac[b], lb, ub, normalize)) <= 1e-9 * max(1.0, abs(omega[b]))
    assert omega[3] <= 1e-9 or k == 1
    assert iters.max() <= 50 * (n + k) + 100


def test_host_mirror_and_critical_point(engine):
    """surrogate._steepest_descent_direction: same signature and return convention as descent.jl:91-135."""
    x = np.array([0.2, 0.7, 0.5])
    J = np.array([[1.0, 0.0, -2.0], [0.5, 1.0, 0.0]])
    d, om = mb._steepest_descent_direction(x, J, np.zeros(3), np.ones(3))
    dr, omr = D.lp_highs(x, J, np.zeros(3), np.ones(3))
    assert abs(om - omr) <= 1e-12 and np.allclose(d, dr, atol=1e-9)        # unique vertex here
    d0, om0 = mb._steepest_descent_direction(x, np.zeros((2, 3)), np.zeros(3), np.ones(3))
    assert np.all(d0 == 0) and om0 == 0.0
    with pytest.raises(NotImplementedError):
        mb._steepest_descent_direction(x, J, np.zeros(3), np.ones(3), A_eq=np.ones((1, 3)), b_eq=np.ones(1))


def test_descent_then_backtrack_on_a_built_model(engine):
    """Jacobian from the device model -> LP -> Armijo batch: one steepest-descent step of descent.jl:187-260 for a batch."""
    from morbit_jl_b200 import synthetic
    rng = np.random.default_rng(5)
    B, n, N = 6, 8, 30
    cfg = mb.RbfConfig(kernel="cubic")
    sites = rng.random((B, N, n)); vals = synthetic.zdt3(sites)
    model, status = engine.build(cfg, sites, vals, [N] * B)
    x = sites[:, 0, :].copy()
    _, J = engine.eval(model, x[:, None, :], False, True)
    d, omega, _, st = engine.descent_direction(J[:, 0], x, np.zeros(n), np.ones(n))
    assert np.all(st == 0) and np.all(omega >= -1e-12)
    nrm = np.abs(d).max(axis=1, keepdims=True)
    xp, mxp, step, idx, mx = engine.backtrack(model, x, d / np.maximum(nrm, 1e-300), nrm[:, 0], omega)
    for b in range(B):
        if omega[b] > 1e-8:
            assert np.all(mx[b] - mxp[b] >= -1e-12)        # a descent step for every output of the model
        assert np.all(xp[b] >= -1e-12) and np.all(xp[b] <= 1 + 1e-12)
    model.free()


def test_backtrack_dev_equals_host_entry_point(engine):
    """mrbf_backtrack_dev (device pointers, no synchronisation) returns exactly what mrbf_backtrack returns."""
    import torch
    from morbit_jl_b200 import synthetic
    rng = np.random.default_rng(11)
    B, n, N = 9, 6, 25
    cfg = mb.RbfConfig(kernel="multiquadric")
    sites = rng.random((B, N, n)); vals = synthetic.zdt3(sites)
    model, _ = engine.build(cfg, sites, vals, [N] * B)
    x = sites[:, 0, :].copy()
    d = rng.normal(size=(B, n)); d /= np.abs(d).max(axis=1, keepdims=True)
    step0 = np.full(B, 0.3); omega = rng.random(B)
    xp, mxp, step, idx, mx = engine.backtrack(model, x, d, step0, omega)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out = engine.backtrack_dev(model, t(x), t(d), t(step0), t(omega))
    engine.sync()
    assert np.array_equal(out[0].cpu().numpy(), idx)
    assert np.array_equal(out[2].cpu().numpy(), xp) and np.array_equal(out[3].cpu().numpy(), mx) and np.array_equal(out[4].cpu().numpy(), mxp)
    model.free()
