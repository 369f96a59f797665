"""Round-2 parity tests (through the C ABI, against the CPU oracle): the shapes the benchmark numbers are quoted on (BASELINE
configs C3 and C5 at full size), thin-plate splines, RbfConfig(optimized_sampling=false), the `Δ ≈ Δ_max` tolerance and the
ownership / error contract of the build entry points."""
import numpy as np
import pytest

import morbit_jl_b200 as mb
from morbit_jl_b200 import synthetic
from oracle import c_oracle as CO
from oracle import rbf_oracle as O
from helpers import random_instances, assert_select_equal

pytestmark = pytest.mark.gpu
RTOL = 1e-10          # north-star tolerance for values / Jacobians (relative to the largest magnitude)


# ------------------------------------------------------------------------------------------------ C5 (eval sweep shape)
@pytest.mark.parametrize("kernel,shape", [("gaussian", 1.0), ("cubic", float("nan"))])
@pytest.mark.parametrize("k", [1, 2])
def test_c5_shape_values_and_jacobians(engine, kernel, shape, k):
    """BASELINE config C5 exactly as bench.py quotes it: 512 centres uniform in [0,1]^50 (seed 0), values [sum x^2, sum sin x][:k],
    Gaussian alpha = 1 and cubic beta = 3, trial points in [c_bar - 0.2, c_bar + 0.2] (seed 1), M = 10^3, 10^4 and 10^5 -- the
    batch sizes that take eval_small / the tile kernels / the DMMA path.  Values and Jacobians <= 1e-10 relative vs the oracle."""
    centers, vals, X = synthetic.eval_sweep(512, 50, k, 10**5, seed=0)
    cfg = mb.RbfConfig(kernel=kernel, shape_parameter=shape)
    wr, lr, st = CO.build_batched(cfg, centers[None], vals[None], [512])
    assert st[0] == 0
    model, status = engine.build(cfg, centers[None], vals[None], [512])
    assert status[0] == 0
    nthr = min(16, CO.max_threads())
    Yr = CO.eval_points(cfg, centers, wr[0], lr[0], X, nthreads=nthr)
    Jr = CO.jac_points(cfg, centers, wr[0], lr[0], X, nthreads=nthr)
    sy, sj = np.abs(Yr).max(), np.abs(Jr).max()
    for M in (10**3, 10**4, 10**5):
        Y, J = engine.eval(model, X[None, :M], True, True)
        Yv, _ = engine.eval(model, X[None, :M], True, False)            # values-only route (DMMA value kernel)
        assert np.abs(Y[0] - Yr[:M]).max() <= RTOL * sy, (M, np.abs(Y[0] - Yr[:M]).max() / sy)
        assert np.abs(Yv[0] - Yr[:M]).max() <= RTOL * sy, (M, np.abs(Yv[0] - Yr[:M]).max() / sy)
        assert np.abs(J[0] - Jr[:M]).max() <= RTOL * sj, (M, np.abs(J[0] - Jr[:M]).max() / sj)
    # a batch that is not a multiple of any tile size, plus the centres themselves (rho = 0)
    Xo = np.vstack((X[:777], centers[:5]))
    Y, J = engine.eval(model, Xo[None], True, True)
    Yo = CO.eval_points(cfg, centers, wr[0], lr[0], Xo); Jo = CO.jac_points(cfg, centers, wr[0], lr[0], Xo)
    assert np.abs(Y[0] - Yo).max() <= RTOL * sy and np.abs(J[0] - Jo).max() <= RTOL * sj
    assert np.abs(Y[0, -5:] - vals[:5]).max() <= 1e-9 * np.abs(vals).max()           # interpolation at the centres
    model.free()


# ------------------------------------------------------------------------------------------------ C3 (full batch)
def test_c3_full_batch_every_instance_matches_oracle(engine):
    """BASELINE config C3 at full size, the batch bench.py times: 4096 ZDT3 n = 30, k = 2 instances, 128-site snapshots, default
    multiquadric RbfConfig.  Training ids [centre; r1; r2; r3; r4] of ALL 4096 instances equal to the oracle's (one fused
    select + build per instance inside the C port, the same routine the CPU arm of bench.py times), and model values and
    Jacobians at trial points in the trust region <= 1e-10 relative on 512 instances."""
    import torch
    from morbit_jl_b200.multistart import MultistartBuilder, upload_batch
    B, n, n_db, k = 4096, 30, 128, 2
    cfg = mb.RbfConfig(kernel="multiquadric")
    host = synthetic.multistart_batch(B, n=n, n_db=n_db, delta=0.1, delta_max=0.5, func=synthetic.zdt3)
    dev = upload_batch(host, "cuda:0")
    builder = MultistartBuilder(engine, cfg, host["delta_max"])
    model, sel, status = builder.step(dev)
    engine.sync()
    assert int((status != 0).sum().item()) == 0
    r = {a: getattr(sel, a).cpu().numpy() for a in ("r1", "n_r1", "r2", "n_r2", "n_r3", "r4", "n_r4")}
    N, ids, w, lam, st = CO.select_and_build_batched(cfg, host["sites"], host["values"], host["x_index"], host["x"], host["delta"],
                                                     host["delta_max"], host["glb"], host["gub"], False, False, host["max_new"],
                                                     func="zdt3", nthreads=min(32, CO.max_threads()))
    assert np.all(st == 0)
    bad = []
    for b in range(B):
        got = np.concatenate([[host["x_index"][b]], r["r1"][b, :r["n_r1"][b]], r["r2"][b, :r["n_r2"][b]],
                              np.zeros(r["n_r3"][b], np.int64), r["r4"][b, :r["n_r4"][b]]])
        if len(got) != N[b] or not np.array_equal(got, ids[b, :N[b]]):
            bad.append(b)
    assert not bad, f"{len(bad)} of {B} instances differ from the oracle, first: {bad[:8]}"
    sample = np.arange(0, B, 8)                       # 512 instances
    rng = np.random.default_rng(3)
    X = host["x"][:, None, :] + 0.2 * (rng.random((B, 6, n)) - 0.5)
    X = np.clip(X, 0.0, 1.0)
    Y, J = engine.eval(model, X, True, True)
    worst_y = worst_j = 0.0
    for b in sample:
        P = host["sites"][b, ids[b, :N[b]] - 1]
        Yr = CO.eval_points(cfg, P, w[b, :N[b]], lam[b], X[b]); Jr = CO.jac_points(cfg, P, w[b, :N[b]], lam[b], X[b])
        worst_y = max(worst_y, np.abs(Y[b] - Yr).max() / np.abs(Yr).max())
        worst_j = max(worst_j, np.abs(J[b] - Jr).max() / np.abs(Jr).max())
    assert worst_y <= RTOL and worst_j <= RTOL, (worst_y, worst_j)
    model.free()


# ------------------------------------------------------------------------------------------------ thin-plate splines (A13)
@pytest.mark.parametrize("n,N,k,deg", [(2, 12, 2, 1), (5, 30, 1, 1), (5, 30, 3, -1), (10, 66, 2, 0), (30, 61, 2, 1)])
def test_thin_plate_spline_order1_build_and_eval(engine, n, N, k, deg):
    """:thin_plate_spline with shape_parameter = 1 (phi = rho^2 log rho, cpd order 2 => linear tail; a lower polynomial_degree is
    raised to 1, U4).  test/rbf_models.jl:27-30 sweeps this kernel with degrees -1..1."""
    rng = np.random.default_rng(100 * n + N)
    cfg = mb.RbfConfig(kernel="thin_plate_spline", shape_parameter=1.0, polynomial_degree=deg)
    B = 3
    S = rng.random((B, N, n))
    V = np.stack([np.sum(S**2, -1), np.sum(np.sin(3 * S), -1), S[..., 0] * S[..., -1]], -1)[..., :k]
    wr, lr, st = CO.build_batched(cfg, S, V, [N] * B, nthreads=3)
    assert np.all(st == 0)
    model, status = engine.build(cfg, S, V, [N] * B)
    assert np.all(status == 0) and model.degree == 1
    X = np.concatenate((rng.random((B, 70, n)), S[:, :3]), axis=1)              # includes training sites (rho = 0: 0 * log 0 := 0)
    Y, J = engine.eval(model, X, True, True)
    Ys, _ = engine.eval(model, X[:, :5], True, False)                           # eval_small route
    for b in range(B):
        Yr = CO.eval_points(cfg, S[b], wr[b], lr[b], X[b]); Jr = CO.jac_points(cfg, S[b], wr[b], lr[b], X[b])
        cond = O.build_model(S[b], V[b], O.RbfConfig(kernel="thin_plate_spline", shape_parameter=1.0, polynomial_degree=deg)).cond
        tol = max(RTOL, 20 * cond * np.finfo(float).eps)
        assert np.all(np.isfinite(Y[b])) and np.all(np.isfinite(J[b]))
        assert np.abs(Y[b] - Yr).max() <= tol * np.abs(Yr).max(), (b, np.abs(Y[b] - Yr).max() / np.abs(Yr).max(), cond)
        assert np.abs(Ys[b] - Yr[:5]).max() <= tol * np.abs(Yr).max()
        assert np.abs(J[b] - Jr).max() <= tol * np.abs(Jr).max(), (b, np.abs(J[b] - Jr).max() / np.abs(Jr).max(), cond)
    model.free()


@pytest.mark.parametrize("n,n_db,shape", [(3, 40, 1.0), (6, 90, 1.0), (5, 60, float("nan"))])
def test_thin_plate_spline_point_search(engine, n, n_db, shape):
    """Rounds 1-4 with the thin-plate spline (round 4 uses phi with the CONFIGURED polynomial degree, RbfModel.jl:374-375); the default
    order k = 2 selects points too; its model build raises the tail to degree 2 (test_quadratic_tail_for_kernels_of_cpd_order_three),
    order 3 would need a cubic tail (MRBF_EUNSUPPORTED)."""
    rng = np.random.default_rng(n * 1000 + n_db)
    B = 8
    cfg = mb.RbfConfig(kernel="thin_plate_spline", shape_parameter=shape)
    sites, x, glb, gub = random_instances(rng, B, n, n_db, True)
    xi = np.ones(B, np.int32); dl = np.full(B, 0.1)
    ref = CO.select_points_batched(cfg, sites, xi, x, dl, 0.5, glb, gub, False, False, 2**31 - 1, nthreads=4)
    res = engine.select_points(cfg, sites, np.full(B, n_db), xi, x, dl, 0.5, glb, gub, False, False, 2**31 - 1)
    assert np.all(res.status == 0)
    assert_select_equal(res, ref, B)
    if shape != shape:          # default order 2 builds (quadratic tail); order 3 is the first unsupported one
        model, status = engine.build(cfg, sites[:, :30], np.sum(sites[:, :30] ** 2, -1, keepdims=True), [30] * B)
        assert np.all(status == 0)
        model.free()
        with pytest.raises(mb.MrbfError) as ei:
            engine.build(mb.RbfConfig(kernel="thin_plate_spline", shape_parameter=3.0), sites[:, :30], np.zeros((B, 30, 1)), [30] * B)
        assert ei.value.code == mb._lib.MRBF_EUNSUPPORTED


# ------------------------------------------------------------------------------------------------ optimized_sampling = false
def test_optimized_sampling_false_keep_and_build(engine):
    """RbfConfig(optimized_sampling=false): coordinate rebuild, no rounds 1/2/4 (RbfModel.jl:564-569, 588, 647-652).  The *_keep entry
    points must hand back a usable handle (no factorisation: general route) and the fused build must give the oracle's model."""
    import torch
    from morbit_jl_b200.multistart import MultistartBuilder, upload_batch
    B, n, n_db = 7, 6, 40
    cfg = mb.RbfConfig(kernel="cubic", optimized_sampling=False)
    host = synthetic.multistart_batch(B, n=n, n_db=n_db, delta=0.05, func=synthetic.zdt3, local_fraction=0.5)
    ref = CO.select_points_batched(cfg, host["sites"], host["x_index"], host["x"], host["delta"], host["delta_max"], host["glb"],
                                   host["gub"], False, False, host["max_new"])
    assert np.all(ref.n_r1 == 0) and np.all(ref.n_r2 == 0) and np.all(ref.n_r4 == 0) and np.all(ref.n_r3 == n)
    # host twins
    res, prepared = engine.select_points_keep(cfg, host["sites"], host["n_db"], host["x_index"], host["x"], host["delta"],
                                              host["delta_max"], host["glb"], host["gub"])
    assert prepared.handle
    assert_select_equal(res, ref, B)
    r3v = np.zeros((B, n, 2))
    for b in range(B):
        r3v[b] = synthetic.zdt3(res.r3_sites[b, :n])
    model, status = engine.build_prepared(cfg, prepared, host["sites"], host["values"], host["x_index"], res, r3v)
    assert np.all(status == 0)
    # a second pass through the SAME handle (recycled): still the right training set, not a stale one
    host2 = synthetic.multistart_batch(B, n=n, n_db=n_db, delta=0.05, func=synthetic.zdt3, local_fraction=0.5, first_instance=50)
    res2, prepared = engine.select_points_keep(cfg, host2["sites"], host2["n_db"], host2["x_index"], host2["x"], host2["delta"],
                                               host2["delta_max"], host2["glb"], host2["gub"], prepared=prepared)
    r3v2 = np.zeros((B, n, 2))
    for b in range(B):
        r3v2[b] = synthetic.zdt3(res2.r3_sites[b, :n])
    model2, status2 = engine.build_prepared(cfg, prepared, host2["sites"], host2["values"], host2["x_index"], res2, r3v2, recycle=model)
    assert np.all(status2 == 0)
    X = host2["x"][:, None, :] + 0.05 * (np.random.default_rng(0).random((B, 5, n)) - 0.5)
    Y, J = engine.eval(model2, X, True, True)
    for b in range(B):
        P = np.vstack([host2["x"][b][None], res2.r3_sites[b, :n]])
        V = synthetic.zdt3(P)
        wr, lr, st = CO.build_batched(cfg, P[None], V[None], [n + 1])
        Yr = CO.eval_points(cfg, P, wr[0], lr[0], X[b]); Jr = CO.jac_points(cfg, P, wr[0], lr[0], X[b])
        assert np.abs(Y[b] - Yr).max() <= RTOL * np.abs(Yr).max() and np.abs(J[b] - Jr).max() <= RTOL * np.abs(Jr).max()
    # device twins through the multistart builder (fused = keep + build_prepared)
    dev = upload_batch(host, "cuda:0")
    builder = MultistartBuilder(engine, cfg, host["delta_max"])
    builder.select(dev); engine.sync()
    r3_dev = torch.from_numpy(r3v).cuda()
    sel2, prep_dev = engine.select_points_keep_dev(cfg, dev.sites, dev.n_db, dev.x_index, dev.x, dev.delta, host["delta_max"], dev.glb,
                                                   dev.gub, dev.flags_in, dev.max_new)
    mdev, sdev = engine.build_prepared_dev(cfg, prep_dev, dev.sites, dev.values, dev.x_index, sel2, r3_dev)
    engine.sync()
    assert np.all(sdev.cpu().numpy() == 0) and np.all(sel2.n_r3.cpu().numpy() == n)
    Yh, _ = engine.eval(mdev, host["x"][:, None, :], True, False)
    assert np.abs(Yh[:, 0] - synthetic.zdt3(host["x"])).max() <= 1e-9          # interpolation at the centre (test/rbf_models.jl:104)
    model2.free(); mdev.free(); prepared.free(); prep_dev.free()


def test_plugin_mirror_with_optimized_sampling_false(engine):
    """prepare_update_model -> eval_missing! -> update_model through the plugin mirror with optimized_sampling = false."""
    rng = np.random.default_rng(5)
    n = 4
    cfg = mb.RbfConfig(kernel="multiquadric", optimized_sampling=False)
    f = lambda x: np.array([np.sum((x - 0.3) ** 2), np.sum(np.sin(x))])
    db = mb.ArrayDB(n)
    x = rng.random(n)
    xid = db.new_result(x, f(x))
    for _ in range(25):
        s = np.clip(x + 0.2 * (rng.random(n) - 0.5), 0, 1); db.new_result(s, f(s))
    sdb = mb.SuperDB({(1,): db}); it = mb.IterData(x, 0.1, {(1,): xid}); scal = mb.VarScaler(np.zeros(n), np.ones(n))
    ac = mb.AlgoConfig()
    meta = mb.prepare_init_model(cfg, (1,), mb.MopStub(), scal, it, sdb, ac)
    assert meta.round1_indices == [] and meta.round4_indices == [] and len(meta.round3_indices) == n and meta.fully_linear
    db.eval_missing(f)
    mod, meta = mb.init_model(meta, cfg, (1,), mb.MopStub(), scal, it, sdb, ac)
    ids = mb._collect_indices(meta)
    P = np.array([db.get_site(i) for i in ids]); V = np.array([db.get_value(i) for i in ids])
    om = O.build_model(P, V, O.RbfConfig(kernel="multiquadric", optimized_sampling=False))
    xt = x + 0.01
    assert np.abs(mb.eval_models(mod, scal, xt) - om.eval(xt)).max() <= RTOL * np.abs(om.eval(xt)).max()
    assert np.abs(mb.get_jacobian(mod, scal, xt) - om.jac(xt)).max() <= 1e-9 * np.abs(om.jac(xt)).max()
    # a second update reuses the kept handle
    meta = mb.prepare_update_model(mod, meta, cfg, (1,), mb.MopStub(), scal, it, sdb, ac)
    db.eval_missing(f)
    mod, meta = mb.update_model(mod, meta, cfg, (1,), mb.MopStub(), scal, it, sdb, ac)
    assert np.abs(mb.eval_models(mod, scal, x) - f(x)).max() <= 1e-9


# ------------------------------------------------------------------------------------------------ Δ ≈ Δ_max tolerance
def test_isapprox_tolerance_follows_the_algorithm_config_precision(engine):
    """RbfModel.jl:588: round 2 is skipped (and fully_linear set) when Δ ≈ Δ_max.  With the reference's default config Δ_max is a
    Float32 literal, so Julia's isapprox uses rtol = sqrt(eps(Float32)) = 3.45e-4; with an AlgorithmConfig{Float64}, 1.49e-8."""
    rng = np.random.default_rng(8)
    B, n, n_db = 6, 5, 60
    cfg = mb.RbfConfig(kernel="cubic")
    sites, x, glb, gub = random_instances(rng, B, n, n_db, True, spread=1.0)
    sites[:, 3:] = np.clip(x[:, None, :] + 0.9 * (rng.random((B, n_db - 3, n)) - 0.5) * 2, 0, 1)     # mostly outside box 1
    xi = np.ones(B, np.int32)
    dmax = 0.5
    dl = np.full(B, dmax * (1 - 1e-4))            # inside the Float32 tolerance, outside the Float64 one
    try:
        for rtol in (O.ISAPPROX_RTOL_F64, O.ISAPPROX_RTOL_F32):
            CO.set_isapprox_rtol(rtol); engine.set_isapprox_rtol(rtol)
            ref = CO.select_points_batched(cfg, sites, xi, x, dl, dmax, glb, gub, False, False, 2**31 - 1)
            res = engine.select_points(cfg, sites, np.full(B, n_db), xi, x, dl, dmax, glb, gub, False, False, 2**31 - 1)
            assert_select_equal(res, ref, B)
            if rtol == O.ISAPPROX_RTOL_F32:
                assert np.all(res.n_r2 == 0)
        # the literal oracle agrees on which branch is taken
        db = O.ArrayDB()
        for s in sites[0]:
            db.new_result(s, [0.0])
        for rtol in (O.ISAPPROX_RTOL_F64, O.ISAPPROX_RTOL_F32):
            meta = O.prepare_update_model(O.RbfMeta(signature=O.RbfConfig().signature()), O.RbfConfig(kernel="cubic"), db, x[0], 1, dl[0], dmax,
                                          glb, gub, isapprox_rtol=rtol)
            CO.set_isapprox_rtol(rtol)
            ref = CO.select_points_batched(cfg, sites[:1], xi[:1], x[:1], dl[:1], dmax, glb, gub, False, False, 2**31 - 1)
            assert meta.round2_indices == list(ref.r2[0, :ref.n_r2[0]]) and meta.round1_indices == list(ref.r1[0, :ref.n_r1[0]])
    finally:
        CO.set_isapprox_rtol(O.ISAPPROX_RTOL_F64); engine.set_isapprox_rtol(O.ISAPPROX_RTOL_F64)


# ------------------------------------------------------------------------------------------------ ownership / error contract
def test_build_error_contract(engine):
    """include/morbit_rbf.h, ownership rule of mrbf_build*: MRBF_ENUMERIC returns a valid handle and per-instance status; other
    errors release the handle that was passed in (no leak, *model NULL)."""
    import ctypes as C
    rng = np.random.default_rng(1)
    B, n, n_db = 4, 3, 20
    cfg = mb.RbfConfig(kernel="cubic")
    host = synthetic.multistart_batch(B, n=n, n_db=n_db, delta=0.1, func=synthetic.zdt3, local_fraction=0.6)
    host["sites"][1, 5] = host["sites"][1, 4]            # a duplicated site: instance 1's reduced kernel matrix is singular
    host["values"] = synthetic.zdt3(host["sites"])
    res, prepared = engine.select_points_keep(cfg, host["sites"], host["n_db"], host["x_index"], host["x"], host["delta"],
                                              host["delta_max"], host["glb"], host["gub"])
    r3v = np.zeros((B, n, 2))
    for b in range(B):
        r3v[b, :res.n_r3[b]] = synthetic.zdt3(res.r3_sites[b, :res.n_r3[b]])
    model, status = engine.build_prepared(cfg, prepared, host["sites"], host["values"], host["x_index"], res, r3v, raise_on_failure=False)
    assert model.handle and status[0] == 0      # (round 4 rejects an exact duplicate itself: tau^2 = 0, so status[1] may well be 0)
    Y, _ = engine.eval(model, host["x"][:, None, :], True, False)
    assert np.abs(Y[0, 0] - synthetic.zdt3(host["x"][0])).max() <= 1e-9
    # an error that is not numerical: kernel mismatch between the kept factorisation and the build -> handle released, NULL back
    handle = C.c_void_p(model.handle); model.handle = None
    bad = mb.to_c_cfg(mb.RbfConfig(kernel="gaussian"))
    st = np.zeros(B, np.int32)
    rc = engine.lib.mrbf_build_prepared(engine.ctx, C.byref(bad), prepared.handle, 2, host["sites"].ctypes.data, host["values"].ctypes.data,
                                        res.r3_sites.ctypes.data, r3v.ctypes.data, host["x_index"].ctypes.data, res.r1.ctypes.data,
                                        res.n_r1.ctypes.data, res.r2.ctypes.data, res.n_r2.ctypes.data, res.n_r3.ctypes.data,
                                        C.byref(handle), st.ctypes.data)
    assert rc == mb._lib.MRBF_EINVAL and not handle.value
    assert b"kernel/shape" in engine.lib.mrbf_last_error(engine.ctx)
    prepared.free()


def test_build_routes_agree_and_fall_back_per_instance(engine):
    """mrbf_build picks its route per instance: the reduced system on the first p training points, or the QR-based kernel
    when those points are not poised / the set is too small or too large.  One batch with all the cases; every instance must match the
    oracle, and the forced QR route (MRBF_BUILD_GENERAL=1) must agree with the default one."""
    import os
    rng = np.random.default_rng(77)
    n, k, ts = 4, 2, 150
    cfg = mb.RbfConfig(kernel="cubic")
    Ns = np.array([40, 5, 3, 140, 40, 12], np.int32)            # 5 = p (no reduced unknowns), 3 < p, 140 - p > 128 (too large)
    B = len(Ns)
    S = rng.random((B, ts, n)); 
    S[4, :5] = np.linspace(0.1, 0.9, 5)[:, None] * np.ones(n)   # first p points collinear: Pi_0 singular -> QR route
    S[5, :5] = S[4, :5] + 1e-7 * rng.standard_normal((5, n))     # first p points affinely dependent up to 1e-7 (well separated): pivot guard -> QR route
    V = np.stack([np.sum(S ** 2, -1), np.sum(np.sin(3 * S), -1)], -1)
    X = rng.random((B, 9, n))
    out = {}
    for env in ("0", "1"):
        os.environ["MRBF_BUILD_GENERAL"] = env
        try:
            model, status = engine.build(cfg, S, V, Ns)
            assert np.all(status == 0), status
            out[env] = engine.eval(model, X, True, True)
            model.free()
        finally:
            os.environ["MRBF_BUILD_GENERAL"] = "0"
    wr, lr, st = CO.build_batched(cfg, S, V, Ns, nthreads=2)
    for b in range(B):
        N = int(Ns[b])
        Yr = CO.eval_points(cfg, S[b, :N], wr[b, :N], lr[b], X[b]); Jr = CO.jac_points(cfg, S[b, :N], wr[b, :N], lr[b], X[b])
        for env in ("0", "1"):
            Y, J = out[env]
            assert np.abs(Y[b] - Yr).max() <= RTOL * np.abs(Yr).max(), (b, env, np.abs(Y[b] - Yr).max() / np.abs(Yr).max())
            assert np.abs(J[b] - Jr).max() <= 1e-9 * np.abs(Jr).max(), (b, env, np.abs(J[b] - Jr).max() / np.abs(Jr).max())


@pytest.mark.parametrize("n,n_db,kernel", [(3, 40, "gaussian"), (6, 100, "inv_multiquadric"), (10, 128, "multiquadric")])
def test_round4_register_kernels_with_a_constant_tail(engine, n, n_db, kernel):
    """polynomial_degree = 0 and no budget for round 3: the found set is the centre alone, which is exactly poised for the constant tail
    (p = 1), so round 4 runs on the register-tiled kernels with a 1 x 1 Pi_0 and a 1 x 1 leverage matrix (tensor-path leverage warp with
    31 rows of padding).  Indices as the oracle's; the model built from the kept factorisation interpolates."""
    rng = np.random.default_rng(n * 1000 + n_db)
    cfg = mb.RbfConfig(kernel=kernel, polynomial_degree=0)
    B = 5
    sites, x, glb, gub = random_instances(rng, B, n, n_db, True, spread=0.6)
    xi = np.ones(B, np.int32); dl = np.full(B, 0.2)
    ref = CO.select_points_batched(cfg, sites, xi, x, dl, 0.5, glb, gub, False, False, 0, nthreads=2)
    res, prep = engine.select_points_keep(cfg, sites, np.full(B, n_db), xi, x, dl, 0.5, glb, gub, False, False, 0)
    assert np.all(res.status == 0)
    for b in range(B):
        for nm, cnt in (("r1", "n_r1"), ("r2", "n_r2"), ("r4", "n_r4")):
            assert list(getattr(res, nm)[b, :getattr(res, cnt)[b]]) == list(getattr(ref, nm)[b, :getattr(ref, cnt)[b]]), (b, nm)
        assert res.n_r3[b] == ref.n_r3[b]
    V = np.stack([np.sum(sites ** 2, -1), np.sum(np.sin(3 * sites), -1)], -1)
    model, status = engine.build_prepared(cfg, prep, sites, V, xi, res, np.zeros((B, n, 2)))
    assert np.all(status == 0)
    for b in range(B):
        ids = [1] + list(res.r1[b, :res.n_r1[b]]) + list(res.r2[b, :res.n_r2[b]]) + list(res.r4[b, :res.n_r4[b]])
        if res.n_r3[b] == 0:
            P = sites[b, np.array(ids) - 1]
            Y, _ = engine.eval(model, np.repeat(P[None, :4], B, axis=0), True, False)
            assert np.abs(Y[b] - V[b, np.array(ids[:4]) - 1]).max() <= 1e-8 * max(1.0, np.abs(V[b]).max())
    prep.free(); model.free()


def test_multistart_builder_evaluates_new_round3_sites(engine):
    """Sparse databases: round 3 creates new sites (RbfModel.jl:269-307).  MultistartBuilder(func=...) evaluates them between the
    selection and the build from the kept factorisation, so the model interpolates the objective AT the new sites too."""
    import torch
    from morbit_jl_b200.multistart import MultistartBuilder, upload_batch
    B, n, n_db = 12, 6, 4                                       # fewer sites than directions: round 3 has to create some
    cfg = mb.RbfConfig(kernel="cubic")
    host = synthetic.multistart_batch(B, n=n, n_db=n_db, delta=0.1, func=synthetic.zdt3, local_fraction=0.5)
    dev = upload_batch(host, "cuda:0")
    builder = MultistartBuilder(engine, cfg, host["delta_max"], func=synthetic.zdt3)
    model, sel, status = builder.step(dev)
    engine.sync()
    assert np.all(status.cpu().numpy() == 0)
    n_r3 = sel.n_r3.cpu().numpy(); r3 = sel.r3_sites.cpu().numpy()
    assert n_r3.sum() > 0, "corpus creates no new sites"
    X = np.zeros((B, n, n)); X[:] = host["x"][:, None, :]
    for b in range(B):
        X[b, :n_r3[b]] = r3[b, :n_r3[b]]
    Y, _ = engine.eval(model, X, True, False)
    for b in range(B):
        for i in range(n_r3[b]):
            assert np.abs(Y[b, i] - synthetic.zdt3(r3[b, i][None])[0]).max() <= 1e-8, (b, i)
    model.free()


@pytest.mark.parametrize("kernel,shape,n,N", [("thin_plate_spline", float("nan"), 2, 12), ("thin_plate_spline", float("nan"), 5, 40),
                                              ("thin_plate_spline", 2.0, 3, 30), ("cubic", 5.0, 4, 40), ("cubic", 5.0, 2, 6)])
def test_quadratic_tail_for_kernels_of_cpd_order_three(engine, kernel, shape, n, N):
    """The default :thin_plate_spline (k = 2, rho^4 log rho) and cubic with beta = 5 are conditionally positive definite of order 3: the
    tail is raised to degree 2 (assumption U4; the reference's test sweeps the default thin plate spline, test/rbf_models.jl:27-30).
    Build (QR route) + generic evaluation kernel against the oracle, whose degree-2 tail is pinned by SciPy's quintic / degree = 2
    (tests/test_scipy_pin.py).  Values, Jacobians, the Armijo batch and interpolation at the sites."""
    rng = np.random.default_rng(n * 97 + N)
    cfg = mb.RbfConfig(kernel=kernel, shape_parameter=shape)
    B, k = 3, 2
    S = rng.random((B, N, n))
    V = np.stack([np.sum(S ** 2, -1) + S[..., 0], np.sum(np.sin(3 * S), -1)], -1)
    model, status = engine.build(cfg, S, V, [N] * B)
    assert np.all(status == 0)
    X = np.concatenate((rng.random((B, 21, n)), S[:, :3]), axis=1)
    Y, J = engine.eval(model, X, True, True)
    Y2, _ = engine.eval(model, X[:, :5], True, False)          # few points: must not take the one-warp-per-point kernel's linear tail
    for b in range(B):
        om = O.build_model(S[b], V[b], O.RbfConfig(kernel=kernel, shape_parameter=shape))
        assert om.degree == 2
        Yr = np.array([om.eval(x) for x in X[b]]); Jr = np.array([om.jac(x) for x in X[b]])
        tol = max(RTOL, 20 * om.cond * np.finfo(float).eps)
        assert np.abs(Y[b] - Yr).max() <= tol * np.abs(Yr).max(), (b, np.abs(Y[b] - Yr).max() / np.abs(Yr).max(), om.cond)
        assert np.abs(Y2[b] - Yr[:5]).max() <= tol * np.abs(Yr).max()
        assert np.abs(J[b] - Jr).max() <= 10 * tol * np.abs(Jr).max(), (b, np.abs(J[b] - Jr).max() / np.abs(Jr).max(), om.cond)
        assert np.abs(Y[b, -3:] - V[b, :3]).max() <= 1e3 * tol * max(1.0, np.abs(V[b]).max())     # interpolation
    model.free()


def test_dev_entry_points_are_ordered_with_torchs_stream():
    """An Engine created without `stream=` enqueues on the library's private non-blocking stream while torch allocates (and zero-fills)
    the outputs of a `*_dev` call on its own current stream.  The Python layer brackets such calls with event waits
    (engine._StreamOrderedLib, mrbf_get_stream): with torch's stream held busy, the fill of the fresh outputs is still pending when the
    call is made -- the results must nevertheless be the ones of the host-buffer path, and readable from torch's stream right after."""
    import torch
    from morbit_jl_b200.multistart import upload_batch
    eng = mb.Engine(0)
    assert eng._stream != torch.cuda.current_stream().cuda_stream
    cfg = mb.RbfConfig(kernel="multiquadric")
    host = synthetic.multistart_batch(64, n=12, n_db=70, func=synthetic.zdt3, local_fraction=0.5)
    ref = eng.select_points(cfg, host["sites"], host["n_db"], host["x_index"], host["x"], host["delta"], host["delta_max"],
                            host["glb"], host["gub"])
    dev = upload_batch(host, "cuda:0")
    torch.cuda.synchronize()
    torch.cuda._sleep(400_000_000)                       # ~0.2 s of work in front of everything torch enqueues from here on
    sel = eng.select_points_dev(cfg, dev.sites, dev.n_db, dev.x_index, dev.x, dev.delta, host["delta_max"], dev.glb, dev.gub,
                                dev.flags_in, dev.max_new)
    got = [t.cpu().numpy() for t in (sel.n_r1, sel.n_r2, sel.n_r3, sel.n_r4, sel.r4)]        # no eng.sync(): torch's stream waits
    for a, b in zip(got[:4], (ref.n_r1, ref.n_r2, ref.n_r3, ref.n_r4)):
        assert np.array_equal(a, b)
    assert ref.n_r4.sum() > 0
    for b in range(64):
        assert np.array_equal(got[4][b, :ref.n_r4[b]], ref.r4[b, :ref.n_r4[b]])
    eng.close()


def test_concurrent_contexts_with_different_shapes_share_the_kernels():
    """Two host threads, each with its own context, launch the SAME kernel instantiations with very different dynamic shared-memory
    sizes (n = 6 vs n = 30: a few KB vs ~100 KB for the round-4 elimination).  The per-function shared-memory limit is process-wide;
    the library only ever raises it (raise_dyn_smem, mrbf_api.cu), so neither thread can see its launch refused -- or served with
    another thread's smaller limit -- in the window between the other thread's set and launch.  Results equal the serial ones."""
    import threading
    shapes = {0: dict(n=6, n_db=40), 1: dict(n=30, n_db=128)}

    def run(eng, which, rep):
        cfg = mb.RbfConfig(kernel="multiquadric")
        h = synthetic.multistart_batch(8, func=synthetic.zdt3, first_instance=10 * rep, **shapes[which])
        res = eng.select_points(cfg, h["sites"], h["n_db"], h["x_index"], h["x"], h["delta"], h["delta_max"], h["glb"], h["gub"])
        return np.concatenate([res.n_r2, res.n_r4] + [res.r4[b, :res.n_r4[b]] for b in range(8)])

    eng = mb.Engine(0)
    serial = {(w, r): run(eng, w, r) for w in shapes for r in range(3)}
    eng.close()
    errors, out = [], {}

    def work(which):
        try:
            e = mb.Engine(0)
            for it in range(40):
                out[(which, it)] = run(e, which, it % 3)
            e.close()
        except Exception as ex:            # an exception in a thread would otherwise only be printed
            errors.append(repr(ex))

    ths = [threading.Thread(target=work, args=(w,)) for w in shapes]
    for t in ths: t.start()
    for t in ths: t.join()
    assert not errors, errors
    for (w, it), v in out.items():
        assert np.array_equal(v, serial[(w, it % 3)]), (w, it)


@pytest.mark.parametrize("kernel", ["multiquadric", "gaussian", "cubic"])
def test_under_poised_round4_with_and_without_hand_over(engine, kernel):
    """Explicit under-poised found sets through mrbf_round4 on the instances of tests/test_handover_property.py: kernels of cpd order <= 1
    are handed over from the literal kernel's prefix run to the register kernels once they are poised, cubic (order 2: the reference goes
    on rejecting what a fresh walk would accept) stays on the literal kernel -- either way the ids are the oracle's."""
    import test_handover_property as H
    for n, n_db, n_found in [(3, 40, 2), (5, 60, 1), (8, 90, 3), (12, 120, 1)]:
        rng = np.random.default_rng(100 * n + n_db + n_found)
        for rep in range(4):
            sites, lb2, ub2 = H._instance(rng, n, n_db)
            found0 = H._found0(rng, sites, lb2, ub2, n_found)
            cfg = mb.RbfConfig(kernel=kernel)
            ref, _ = CO.round4(O.RbfConfig(kernel=kernel), sites, lb2, ub2, found0)
            r4, n_r4, status = engine.round4(cfg, sites[None], [n_db], lb2[None], ub2[None], np.array([found0]), [len(found0)])
            assert status[0] == 0 and [int(v) for v in r4[0, :n_r4[0]]] == [int(v) for v in ref], (n, rep, found0)
