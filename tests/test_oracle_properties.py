"""The properties the reference's own test (test/rbf_models.jl) pins, checked on the oracle: this is the only
pinning the reference offers (it holds no golden vectors), see SURVEY.md §4 and §8(c)."""
import numpy as np
import pytest

from oracle import rbf_oracle as O

f1 = lambda x: np.array([np.sum(x**2)])      # test/rbf_models.jl:4


def _init(cfg, n, constrained, rng, algo_max_evals=O.INT_MAX):
    glb = np.full(n, 0.25 if constrained else -np.inf)
    gub = np.full(n, 0.75 if constrained else np.inf)
    # Morbit scales a fully boxed problem to the unit cube (VarScaler.jl:205-213)
    if constrained:
        glb, gub = np.zeros(n), np.ones(n)
    x0 = rng.random(n)
    db = O.ArrayDB()
    xi = db.new_result(x0, f1(x0))
    meta = O.RbfMeta(signature=cfg.signature())
    O.prepare_update_model(meta, cfg, db, x0, xi, float(np.float32(0.1)), float(np.float32(0.5)), glb, gub,
                           ensure_fully_linear=True, num_objf_evals=1, algo_max_evals=algo_max_evals)
    n_evals = 1 + db.eval_missing(f1)
    return meta, db, x0, xi, glb, gub, n_evals


@pytest.mark.parametrize("n", [2, 5, 10])
@pytest.mark.parametrize("kernel", ["cubic", "inv_multiquadric", "multiquadric", "gaussian"])
@pytest.mark.parametrize("deg", [-1, 0, 1])
@pytest.mark.parametrize("constrained", [True, False])
def test_rbf_models_properties(n, kernel, deg, constrained):
    rng = np.random.default_rng(1234 + n)
    # :35-44  max_evals = 1 -> exactly one evaluation, model builds from a single point
    cfg = O.RbfConfig(kernel=kernel, polynomial_degree=deg, max_evals=1, max_model_points=1)
    meta, db, x0, xi, glb, gub, n_evals = _init(cfg, n, constrained, rng)
    assert n_evals == 1
    mod = O.update_model(meta, cfg, db)
    assert np.allclose(mod.eval(x0), f1(x0))
    # :47-59  many unevaluated sites in the local box -> fully linear after the update (degree 1)
    if deg == 1:
        lb, ub = O.local_bounds(x0, 0.1, glb, gub)
        for _ in range(50 * n):
            db.new_result(lb + (ub - lb) * rng.random(n), None)
        meta2 = O.RbfMeta(signature=cfg.signature())
        O.prepare_update_model(meta2, cfg, db, x0, xi, 0.1, 0.5, glb, gub, ensure_fully_linear=True)
        assert meta2.fully_linear
    # :67-71  budget through the algorithm config
    cfg = O.RbfConfig(kernel=kernel, polynomial_degree=deg)
    meta, db, x0, xi, glb, gub, n_evals = _init(cfg, n, constrained, rng, algo_max_evals=1)
    assert n_evals == 1
    # :74-86  round 4 runs with only the centre as found index
    lb2, ub2 = O.local_bounds(x0, cfg.theta_enlarge_2 * 0.5, glb, gub)
    for _ in range(10 * n):
        db.new_result(lb2 + (ub2 - lb2) * rng.random(n), None)
    O.rbf_round4(db, lb2, ub2, x0, 0.1, [xi], cfg)
    # :89-96  default config: fully linear at init
    if deg == 1:
        meta, db, x0, xi, glb, gub, n_evals = _init(cfg, n, constrained, rng)
        assert meta.fully_linear and n_evals == n + 1
    else:
        db.eval_missing(f1)
    mod = O.update_model(meta, cfg, db) if deg == 1 else O.update_model(O.RbfMeta(center_index=xi), cfg, db)
    # :104 interpolation at the centre; :105-111 gradient == jacobian row, ~ finite differences
    assert np.allclose(mod.eval(x0)[-1], f1(x0)[0])
    g = mod.grad(x0, 1)
    assert np.array_equal(g, mod.jac(x0)[0])
    h = 1e-6
    fd = np.array([(mod.eval(x0 + h * e)[0] - mod.eval(x0 - h * e)[0]) / (2 * h) for e in np.eye(n)])
    assert np.allclose(g, fd, rtol=1e-4, atol=1e-5)


def test_rounds_1_to_3_shared_between_kernels():
    """test/rbf_models.jl:123-162."""
    rng = np.random.default_rng(7)
    n = 2
    cfg1, cfg2 = O.RbfConfig(kernel="gaussian"), O.RbfConfig(kernel="multiquadric")
    x0 = rng.random(n)
    db1, db2 = O.ArrayDB(), O.ArrayDB()
    i1 = db1.new_result(x0, [0.0]); i2 = db2.new_result(x0, [0.0])
    for _ in range(20):
        xi = rng.random(n)
        db1.new_result(xi, None); db2.new_result(xi, None)
    glb, gub = np.full(n, -np.inf), np.full(n, np.inf)
    m1 = O.RbfMeta(signature=cfg1.signature()); m2 = O.RbfMeta(signature=cfg2.signature())
    O.prepare_update_model(m1, cfg1, db1, x0, i1, 0.1, 0.5, glb, gub, ensure_fully_linear=True, meta_array=[])
    O.prepare_update_model(m2, cfg2, db2, x0, i2, 0.1, 0.5, glb, gub, ensure_fully_linear=True, meta_array=[(m1, db1)])
    for fn in ("round1_indices", "round2_indices", "round3_indices"):
        a, b = getattr(m1, fn), getattr(m2, fn)
        assert len(a) == len(b)
        assert all(np.array_equal(db1.get_site(i), db2.get_site(j)) for i, j in zip(a, b))


def test_backtrack_batched_equals_sequential():
    """evaluating all step sizes at once and taking the first passing index equals the loop."""
    rng = np.random.default_rng(3)
    n = 4
    cfg = O.RbfConfig(kernel="cubic")
    S = rng.random((15, n)); V = np.stack([np.sum((S - 0.3)**2, 1), np.sum((S + 0.2)**2, 1)], 1)
    mod = O.build_model(S, V, cfg)
    x = rng.random(n); d = -mod.jac(x).sum(0); d /= np.abs(d).max()
    xp, mxp, step, i = O.backtrack(mod.eval, x, d, 1.0, 0.5)
    sig = 1.0
    sigs = []
    for _ in range(118):
        sigs.append(sig); sig *= 0.75
    mx = mod.eval(x)
    ok = [bool(np.all(mx - mod.eval(x + s * d) >= s * 1e-6 * 0.5)) for s in sigs]
    assert i == ok.index(True)
    assert np.array_equal(xp, x + sigs[i] * d)
