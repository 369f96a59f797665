"""Lock-step batched `optimize` with a device-resident database (morbit.jl_b200/lockstep.py) against
the scalar restatement of iterate! (oracle/iterate_oracle.py): same databases, iterates, radii, classifications and stop codes."""
import math

import numpy as np
import pytest

import morbit_jl_b200 as mb
from morbit_jl_b200 import lockstep as L, synthetic
from oracle import rbf_oracle as O, iterate_oracle as IO

pytestmark = pytest.mark.gpu


def test_db_append_and_model_scatter(engine):
    """mrbf_db_append_dev = new_result! (Databases.jl:174-183) incl. the capacity guard; mrbf_model_scatter_dev swaps instances."""
    import torch
    dev = "cuda:0"
    rng = np.random.default_rng(5)
    B, n, k, cap = 7, 4, 2, 6
    sites = torch.zeros((B, cap, n), dtype=torch.float64, device=dev); values = torch.full((B, cap, k), math.nan, dtype=torch.float64, device=dev)
    n_db = torch.tensor([0, 1, 2, 3, 4, 5, 6], dtype=torch.int32, device=dev)
    new_s = torch.from_numpy(rng.random((B, 3, n))).to(dev); new_v = torch.from_numpy(rng.random((B, 3, k))).to(dev)
    n_add = torch.tensor([3, 2, 0, 3, 3, 1, 1], dtype=torch.int32, device=dev)
    first, status = engine.db_append_dev(sites, values, n_db, new_s, new_v, n_add)
    engine.sync()
    assert status.cpu().tolist() == [0, 0, 0, 0, 1, 0, 1]                    # 4 + 3 > 6 and 6 + 1 > 6: refused, nothing written
    assert n_db.cpu().tolist() == [3, 3, 2, 6, 4, 6, 6]
    assert first.cpu().tolist() == [1, 2, 3, 4, 0, 6, 0]
    s, v = sites.cpu().numpy(), values.cpu().numpy()
    np.testing.assert_array_equal(s[0, :3], new_s.cpu().numpy()[0]); np.testing.assert_array_equal(v[3, 3:6], new_v.cpu().numpy()[3])
    np.testing.assert_array_equal(s[1, 1:3], new_s.cpu().numpy()[1, :2]); assert np.all(s[4] == 0) and np.all(np.isnan(v[4]))
    first, status = engine.db_append_dev(sites, values, n_db, new_s, None, torch.tensor([1, 0, 1, 0, 0, 0, 0], dtype=torch.int32, device=dev))
    engine.sync()
    assert np.all(np.isnan(values.cpu().numpy()[0, 3])) and n_db.cpu().tolist()[:3] == [4, 3, 3]        # value-less result: NaN row
    # model scatter
    cfg = mb.RbfConfig(kernel="cubic")
    P = rng.random((5, 8, 3)); V = rng.random((5, 8, 2))
    big, _ = engine.build(cfg, P, V, [8] * 5)
    small, _ = engine.build(cfg, P[[4, 0]] + 0.5, V[[4, 0]], [8, 7])
    engine.model_scatter_dev(big, small, torch.tensor([3, -1], dtype=torch.int32, device=dev))
    X = rng.random((5, 2, 3))
    Y, _ = engine.eval(big, X)
    ref, _ = engine.build(cfg, np.concatenate([P[:3], P[4:5] + 0.5, P[4:]]), np.concatenate([V[:3], V[4:5], V[4:]]), [8] * 5)
    Yr, _ = engine.eval(ref, X)
    np.testing.assert_array_equal(Y, Yr)


class _Feed:
    """Hands the oracle the directions the GPU path computed for instance b, in call order (see oracle/iterate_oracle.py)."""
    def __init__(self, calls, b):
        self.calls = [(c["d"][b], float(c["omega"][b])) for c in calls if c["mask"][b]]
        self.i = 0

    def __call__(self, run, J):
        d, om = self.calls[self.i]
        self.i += 1
        return d, om


def _compare(drv, func, x0, glb, gub, ocfg, oac, min_exact):
    """Replays every instance through the oracle; returns the number of (instance, iteration) states compared."""
    B = x0.shape[0]
    compared = walls = reordered = 0
    for b in range(B):
        run = IO.Run(func, x0[b], glb, gub, ocfg, oac, direction=_Feed(drv.lp_calls, b))
        for t, tr in enumerate(drv.trace):
            if run.ret_code != IO.CONTINUE:
                assert tr["ret"][b] != L.CONTINUE
                break
            r = run.iterate()
            if r.knife:
                break                                  # a pivot / tau^2 test decided by less than 1e-9 relative: stop comparing here
            ctx = (b, t, r, {k_: v[b] for k_, v in tr.items() if k_ not in ("iter_counter", "r1", "r2", "r4")})
            knife = (abs(r.rho - oac.nu_success) < 1e-6 or abs(r.rho - oac.nu_accept) < 1e-6) if math.isfinite(r.rho) else False
            if knife and (tr["it_stat"][b] != r.it_stat):
                break                                  # acceptance ratio within 1e-6 of a threshold: either classification is legitimate
            on_wall = bool(np.any((run.x == glb) | (run.x == gub)))
            if on_wall and (tr["n_db"][b] != r.n_db or tr["it_stat"][b] != r.it_stat):
                # iterate exactly on a wall of the box (ZDT3 converges into the corner x = 0): shifted sites then have exact zeros,
                # the Householder pivots of qr(Y) (AffinelyIndependentPoints.jl:4-11) are rounding noise (+-1e-16), their sign picks
                # the basis Z, and the reference's score ||Z Z' s||_inf depends on that basis (columns scaled by their inf-norm):
                # which points round 1 keeps is decided by noise in the reference itself.  Seen once in this corpus.
                walls += 1
                break
            assert tr["ret"][b] == r.ret_code and tr["it_stat"][b] == r.it_stat, ctx
            assert tr["x_index"][b] == r.x_index and tr["n_db"][b] == r.n_db and tr["num_evals"][b] == r.num_evals, ctx
            assert bool(tr["fully_linear"][b]) == r.fully_linear, ctx
            assert abs(tr["delta"][b] - r.delta) <= 1e-12 * r.delta, ctx
            np.testing.assert_allclose(tr["x"][b], r.x, rtol=0, atol=1e-9, err_msg=str(ctx))
            np.testing.assert_allclose(tr["fx"][b], r.fx, rtol=1e-9, atol=1e-12, err_msg=str(ctx))
            if math.isfinite(r.rho):
                assert abs(tr["rho"][b] - r.rho) <= 1e-6 * max(1.0, abs(r.rho)), ctx
            ids = ([int(tr["center"][b])] + list(tr["r1"][b][: tr["n_r1"][b]]) + list(tr["r2"][b][: tr["n_r2"][b]])
                   + [int(tr["r3_first"][b]) + i for i in range(tr["n_r3"][b])] + list(tr["r4"][b][: tr["n_r4"][b]]))
            ids = [int(v) for v in ids]
            if ids != r.training_ids and sorted(ids) == sorted(r.training_ids):
                # same training SET, different greedy order: the databases of this algorithm are exactly structured (coordinate
                # steps from x0, LP-vertex steps with |d_j| = 1), candidate scores and Householder pivots tie exactly, and the
                # order is decided by rounding noise in the reference as well -- the NumPy and the C oracle disagree with each
                # other on these states, while any 1e-15 perturbation of the sites makes all implementations agree again
                # (tests/test_gpu_plugin.py tolerates the same).  The states diverge legitimately from here.
                reordered += 1
                compared += 1
                break
            assert ids == r.training_ids, ctx
            compared += 1
    assert compared >= min_exact and walls <= 1, (compared, walls)
    return compared


def test_lockstep_two_parabolas_matches_oracle():
    """BASELINE config C1 as a batch: 12 starting points, n = 2, cubic, unbounded; 25 lock-step iterations."""
    rng = np.random.default_rng(1)
    B, n = 12, 2
    x0 = np.vstack([[-np.pi, 2.71828], rng.uniform(-3, 3, (B - 1, n))])
    glb, gub = np.full(n, -np.inf), np.full(n, np.inf)
    ac = L.AlgorithmConfig(max_iter=25)
    drv = L.LockstepDriver(mb.RbfConfig(kernel="cubic"), synthetic.two_parabolas, x0, glb, gub, ac, capacity=128, record=True)
    x, fx, ret = drv.run()
    assert np.all(ret != L.CONTINUE) and np.all(ret != L.NUMERIC) and np.all(ret != L.DB_FULL)
    assert np.all(drv.n_db.cpu().numpy() == drv.num_evals.cpu().numpy())          # every evaluated site is a database row
    assert np.all(np.abs(x[:, 0] - x[:, 1]) < 0.3) and np.all(np.abs(x) < 1.2)    # near the Pareto set of the two parabolas: x1 = x2 in [-1, 1]
    _compare(drv, lambda z: synthetic.two_parabolas(np.asarray(z)), x0, glb, gub, O.RbfConfig(kernel="cubic"), IO.AlgoConfig(max_iter=25),
             min_exact=B * 10)


def test_lockstep_zdt3_n30_matches_oracle():
    """BASELINE config C3 (shape): ZDT3 n = 30, k = 2, multiquadric, Halton starting points; 6 lock-step iterations of 6 instances
    replayed through the NumPy oracle."""
    B, n = 6, 30
    x0 = synthetic.halton(B, n)
    glb, gub = np.zeros(n), np.ones(n)
    ac = L.AlgorithmConfig(max_iter=6)
    drv = L.LockstepDriver(mb.RbfConfig(kernel="multiquadric"), synthetic.zdt3, x0, glb, gub, ac, capacity=128, record=True)
    x, fx, ret = drv.run()
    assert np.all(ret != L.CONTINUE) and np.all(ret != L.NUMERIC)
    _compare(drv, lambda z: synthetic.zdt3(np.asarray(z)), x0, glb, gub, O.RbfConfig(kernel="multiquadric"), IO.AlgoConfig(max_iter=6),
             min_exact=B * 4)


def test_lockstep_c3_batch_properties():
    """512 ZDT3 instances to completion (max_iter = 30): valid stop codes, database bookkeeping, monotone objectives
    (strict acceptance test with nu_accept = 0: an accepted step never increases any objective)."""
    B, n = 512, 30
    x0 = synthetic.halton(B, n)
    f0 = synthetic.zdt3(x0)
    drv = L.LockstepDriver(mb.RbfConfig(kernel="multiquadric"), synthetic.zdt3, x0, np.zeros(n), np.ones(n), L.AlgorithmConfig(max_iter=30),
                           capacity=128)
    x, fx, ret = drv.run()
    assert np.all(np.isin(ret, [L.MAX_ITER, L.CRITICAL, L.TOLERANCE, L.DB_FULL, L.BUDGET_EXHAUSTED]))
    assert np.mean(ret == L.DB_FULL) < 0.25
    ok = ret != L.DB_FULL                                   # a refused append (capacity) is the one case where an evaluation has no row
    assert np.all(drv.n_db.cpu().numpy()[ok] == drv.num_evals.cpu().numpy()[ok])
    assert np.all(fx <= f0 + 1e-12) and np.mean(np.any(fx < f0 - 1e-6, axis=1)) > 0.6 and fx[:, 1].mean() < 0.75 * f0[:, 1].mean()
    np.testing.assert_allclose(fx, synthetic.zdt3(x), rtol=0, atol=1e-14)
    assert np.all((x >= 0) & (x <= 1))
    # the iterate is a row of its own database with its own values
    xi = drv.x_index.cpu().numpy().astype(int) - 1
    s = drv.sites.cpu().numpy()[np.arange(B), xi]; v = drv.values.cpu().numpy()[np.arange(B), xi]
    np.testing.assert_array_equal(s, x); np.testing.assert_array_equal(v, fx)
