"""Pins the fast C restatement (oracle/rbf_oracle.c) against the literal NumPy/LAPACK restatement
(oracle/rbf_oracle.py) -- the two were written independently from the reference sources."""
import zlib

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import rbf_oracle as O


def _db(rng, n, n_db, x0, glb, gub, spread=0.6):
    db = O.ArrayDB()
    xi = db.new_result(x0, [0.0])
    for _ in range(n_db - 1):
        db.new_result(np.clip(x0 + (rng.random(n) * 2 - 1) * spread * rng.random(), glb, gub), [0.0])
    return db, xi


CASES = [
    # n, kernel, deg, n_db, boxed, ensure_fl, max_new, delta
    (2, "cubic", 1, 40, True, True, 10**6, 0.1),
    (2, "gaussian", 1, 40, False, False, 10**6, 0.1),
    (5, "multiquadric", 1, 120, True, False, 10**6, 0.1),
    (5, "cubic", 1, 40, True, False, 2, 0.1),          # budget-limited round 3 -> N < n+1 entering round 4
    (5, "gaussian", 1, 40, True, False, 0, 0.1),
    (6, "multiquadric", 0, 50, True, True, 1, 0.1),
    (4, "gaussian", -1, 30, True, True, 10, 0.1),
    (5, "cubic", 1, 12, True, True, 10**6, 0.1),        # few sites: round 3 samples along improving directions
    (8, "cubic", 1, 3, True, False, 3, 0.05),
    (10, "inv_multiquadric", 1, 300, False, True, 10**6, 0.1),
    (30, "multiquadric", 1, 128, True, False, 10**6, 0.1),
]


@pytest.mark.parametrize("n,kernel,deg,n_db,boxed,efl,max_new,delta", CASES)
def test_select_points_c_matches_py(n, kernel, deg, n_db, boxed, efl, max_new, delta):
    rng = np.random.default_rng(zlib.crc32(repr((n, kernel, deg, n_db)).encode()))
    cfg = O.RbfConfig(kernel=kernel, polynomial_degree=deg)
    glb = np.full(n, 0.0 if boxed else -np.inf)
    gub = np.full(n, 1.0 if boxed else np.inf)
    x0 = rng.random(n)
    db, xi = _db(rng, n, n_db, x0, glb, gub)
    sites = np.array(db.sites)
    meta = O.RbfMeta(signature=cfg.signature())
    O.prepare_update_model(meta, cfg, db, x0, xi, delta, 0.5, glb, gub, ensure_fully_linear=efl,
                           algo_max_evals=(max_new + 1) if max_new < 10**6 else O.INT_MAX)
    res = CO.select_points_batched(cfg, sites[None], [xi], x0[None], [delta], 0.5, glb, gub, efl, False, max_new)
    assert list(res.r1[0, :res.n_r1[0]]) == meta.round1_indices
    assert list(res.r2[0, :res.n_r2[0]]) == meta.round2_indices
    assert list(res.r4[0, :res.n_r4[0]]) == meta.round4_indices
    assert res.n_r3[0] == len(meta.round3_indices)
    assert bool(res.fully_linear[0]) == meta.fully_linear
    if meta.round3_indices:
        r3 = np.array([db.get_site(i) for i in meta.round3_indices])
        np.testing.assert_allclose(res.r3_sites[0, :res.n_r3[0]], r3, rtol=0, atol=1e-14)
    if res.n_dirs[0]:
        np.testing.assert_allclose(res.dirs[0, :res.n_dirs[0]], np.array(meta.improving_directions), rtol=0, atol=1e-12)


@pytest.mark.parametrize("n,kernel,deg,N,k", [(5, "cubic", 1, 20, 2), (5, "multiquadric", 1, 21, 1), (5, "gaussian", -1, 15, 3),
                                              (5, "inv_multiquadric", 0, 15, 2), (6, "cubic", 1, 3, 2),
                                              (30, "multiquadric", 1, 61, 2), (4, "cubic", 1, 12, 1), (3, "cubic", 1, 1, 1)])
def test_build_eval_c_matches_py(n, kernel, deg, N, k):
    rng = np.random.default_rng(N * 31 + n)
    cfg = O.RbfConfig(kernel=kernel, polynomial_degree=deg, shape_parameter=2.0 if kernel == "gaussian" else float("nan"))
    S = rng.random((N, n))
    V = np.stack([np.sum(S**2, 1), np.sum(np.sin(S), 1), S[:, 0]], 1)[:, :k]
    m = O.build_model(S, V, cfg)
    w, lam, st = CO.build_batched(cfg, S[None], V[None], [N])
    assert st[0] == 0
    X = np.vstack((rng.random((7, n)), S[:2]))
    Ypy = np.array([m.eval(x) for x in X]); Jpy = np.array([m.jac(x) for x in X])
    Yc = CO.eval_points(cfg, S, w[0], lam[0], X); Jc = CO.jac_points(cfg, S, w[0], lam[0], X)
    np.testing.assert_allclose(Yc, Ypy, rtol=1e-10, atol=1e-10 * np.abs(Ypy).max())
    np.testing.assert_allclose(Jc, Jpy, rtol=0, atol=1e-9 * max(1.0, np.abs(Jpy).max()))
    # interpolation at the sites
    np.testing.assert_allclose(CO.eval_points(cfg, S, w[0], lam[0], S), V, rtol=0, atol=1e-8 * max(1.0, np.abs(V).max()))


def test_intersect_box_c_matches_py():
    rng = np.random.default_rng(5)
    for _ in range(200):
        n = int(rng.integers(1, 6))
        lb = rng.random(n) - 1.0; ub = lb + rng.random(n) + 0.1
        x = lb + (ub - lb) * rng.random(n)
        if rng.random() < 0.3:
            x[0] = lb[0]
        if rng.random() < 0.3:
            x[-1] = ub[-1]
        d = rng.standard_normal(n) * (rng.random(n) < 0.8)
        a = O.intersect_box_absmax(x, d, lb, ub); b = CO.intersect_box_absmax(x, d, lb, ub)
        assert a == b or (np.isinf(a) and np.isinf(b))


def test_round4_below_poised_is_rounding_noise():
    """Documented limit of parity (found by the randomised GPU sweep): when round 4 starts with FEWER points than polynomial basis
    functions (budget-limited round 3, N0 < n + 1), the reference's test quantity tau^2 = sigma - ||L^-1 v||^2 (RbfModel.jl:447-452) is
    zero in exact arithmetic for candidates that do not enlarge the span, i.e. +-1e-17 in floating point against a threshold of
    1e-28: its accept / reject decisions are coin flips.  The literal oracle and its C twin -- two correct restatements -- legitimately
    disagree on such an instance; everything decided before round 4 is identical."""
    rng = np.random.default_rng(7)
    n, n_db, max_new = 9, 22, 3
    cfg = O.RbfConfig(kernel="multiquadric")
    glb, gub = np.zeros(n), np.ones(n)
    noise_seen = False
    for rep in range(12):
        x0 = rng.random(n)
        db, xi = _db(rng, n, n_db, x0, glb, gub)
        sites = np.array(db.sites)
        meta = O.RbfMeta(signature=cfg.signature())
        t4 = O.Round4Trace()
        O.prepare_update_model(meta, cfg, db, x0, xi, 0.02, 0.5, glb, gub, ensure_fully_linear=True, algo_max_evals=max_new + 1, trace4=t4)
        res = CO.select_points_batched(cfg, sites[None], [xi], x0[None], [0.02], 0.5, glb, gub, True, False, max_new)
        assert list(res.r1[0, :res.n_r1[0]]) == meta.round1_indices and list(res.r2[0, :res.n_r2[0]]) == meta.round2_indices
        assert res.n_r3[0] == len(meta.round3_indices)
        if 1 + len(meta.round1_indices) + len(meta.round2_indices) + len(meta.round3_indices) < n + 1:
            tiny = [v for v in t4.tau2 if abs(v) < 1e-12]
            noise_seen = noise_seen or len(tiny) > 0
            if list(res.r4[0, :res.n_r4[0]]) != meta.round4_indices:
                assert len(tiny) > 0        # a disagreement is always explained by a noise-level tau^2
    assert noise_seen


def test_under_poised_round4_c_port_vs_literal_per_case_proof():
    """The corpus of tests/test_gpu_parity.py::test_select_points_randomised_sweep, with the C port standing in for the GPU: on every
    instance whose round 4 starts with fewer points than polynomial basis functions the C port either reproduces the literal oracle's
    lists or first departs from them at a candidate whose tau^2 the literal oracle itself computed as an exact cancellation
    (|tau^2| <= 1e-8 of its two terms) -- a coin flip in the reference's own operation order."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import random_instances, literal_round4_verdict
    rng = np.random.default_rng(20261018)
    kernels = ["cubic", "multiquadric", "gaussian", "inv_multiquadric"]
    count = {"equal": 0, "noise": 0}
    for trial in range(48):
        n = int(rng.integers(2, 13)); n_db = int(rng.integers(1, 129)); B = 6
        kernel = kernels[trial % 4]
        mmp = int(rng.choice([-1, -1, n + 2, 2 * n + 1, n + 1 + int(rng.integers(1, 9))]))
        cfg = O.RbfConfig(kernel=kernel, max_model_points=mmp)
        spread = float(rng.choice([0.15, 0.6, 1.2]))
        sites, x, glb, gub = random_instances(rng, B, n, n_db, bool(trial % 3), spread=spread)
        n_dbs = np.minimum(n_db, rng.integers(1, n_db + 1, size=B)).astype(np.int32); n_dbs[0] = n_db
        dl = rng.choice([0.02, 0.1, 0.3], size=B)
        efl = bool(trial % 2)
        max_new = int(rng.choice([0, 1, 3, 2**31 - 1]))
        for b in range(B):
            ref = CO.select_points_batched(cfg, sites[b:b + 1, :n_dbs[b]], [1], x[b:b + 1], dl[b:b + 1], 0.5, glb, gub, efl, False, max_new)
            if 1 + int(ref.n_r1[0]) + int(ref.n_r2[0]) + int(ref.n_r3[0]) >= n + 1:
                continue
            got = dict(r1=ref.r1[0, :ref.n_r1[0]], r2=ref.r2[0, :ref.n_r2[0]], r4=ref.r4[0, :ref.n_r4[0]], n_r3=ref.n_r3[0])
            verdict, info = literal_round4_verdict(cfg, sites[b, :n_dbs[b]], x[b], dl[b], 0.5, glb, gub, efl, max_new, got)
            assert verdict in count, (trial, b, verdict, info)
            count[verdict] += 1
    assert count["equal"] >= 40 and count["noise"] <= 8, count
