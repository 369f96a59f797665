"""Generates tests/golden/*.npz from the literal NumPy/LAPACK oracle (oracle/rbf_oracle.py).

The reference holds no golden vectors for this path and cannot be executed here (no Julia), so these
fixtures pin the ORACLE's outputs, not the reference's; they exist so that the C oracle, the CUDA path and
future rounds are all compared against the same frozen numbers.    python oracle/gen_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import rbf_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

SELECT = [  # name, n, kernel, deg, n_db, boxed, ensure_fully_linear, max_new, delta, max_model_points
    ("sel_n2_cubic", 2, "cubic", 1, 30, False, True, None, 0.1, -1),
    ("sel_n5_mq", 5, "multiquadric", 1, 80, True, False, None, 0.1, -1),
    ("sel_n5_budget2", 5, "cubic", 1, 30, True, False, 2, 0.1, -1),
    ("sel_n6_deg0", 6, "multiquadric", 0, 40, True, True, 1, 0.1, -1),
    ("sel_n10_gauss", 10, "gaussian", 1, 150, True, False, None, 0.07, -1),
    ("sel_n30_mq_cap61", 30, "multiquadric", 1, 100, True, False, None, 0.1, 61),
]
BUILD = [  # name, n, kernel, deg, N, k, shape
    ("mod_n2_cubic", 2, "cubic", 1, 6, 2, float("nan")),
    ("mod_n5_mq", 5, "multiquadric", 1, 21, 2, float("nan")),
    ("mod_n5_gauss_nopoly", 5, "gaussian", -1, 15, 1, 2.0),
    ("mod_n5_imq_deg0", 5, "inv_multiquadric", 0, 15, 2, float("nan")),
    ("mod_n30_cubic", 30, "cubic", 1, 61, 2, float("nan")),
    ("mod_n6_underdetermined", 6, "cubic", 1, 3, 1, float("nan")),
]
TRAJ = [  # name, starting points, max_iter: example_two_parabolas.jl (x0 of the example first)
    ("traj_two_parabolas", [[-np.pi, 2.71828], [2.0, -1.5], [0.4, 2.9], [-2.2, -0.6]], 25),
]
DESCENT = [  # name, n, k, boxed, normalize
    ("lp_n2_k2", 2, 2, False, True),
    ("lp_n30_k2_box", 30, 2, True, True),
    ("lp_n12_k5_box_raw", 12, 5, True, False),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    for i, (name, n, kernel, deg, n_db, boxed, efl, max_new, delta, mmp) in enumerate(SELECT):
        rng = np.random.default_rng(1000 + i)
        cfg = O.RbfConfig(kernel=kernel, polynomial_degree=deg, max_model_points=mmp)
        glb = np.full(n, 0.0 if boxed else -np.inf); gub = np.full(n, 1.0 if boxed else np.inf)
        x = rng.random(n)
        db = O.ArrayDB(); xi = db.new_result(x, [0.0])
        for _ in range(n_db - 1):
            db.new_result(np.clip(x + (rng.random(n) * 2 - 1) * 0.6 * rng.random(), glb, gub), [0.0])
        sites = np.array(db.sites)
        meta = O.RbfMeta(signature=cfg.signature())
        O.prepare_update_model(meta, cfg, db, x, xi, delta, 0.5, glb, gub, ensure_fully_linear=efl,
                               algo_max_evals=O.INT_MAX if max_new is None else max_new + 1)
        r3 = np.array([db.get_site(j) for j in meta.round3_indices]).reshape(-1, n)
        np.savez(os.path.join(OUT, name + ".npz"), kind="select", kernel=kernel, deg=deg, mmp=mmp, sites=sites, x=x, x_index=xi,
                 delta=delta, delta_max=0.5, glb=glb, gub=gub, ensure_fully_linear=efl,
                 max_new=2**31 - 1 if max_new is None else max_new, r1=np.array(meta.round1_indices, np.int32),
                 r2=np.array(meta.round2_indices, np.int32), r3_sites=r3, r4=np.array(meta.round4_indices, np.int32),
                 dirs=np.array(meta.improving_directions).reshape(-1, n), fully_linear=meta.fully_linear)
    for i, (name, n, kernel, deg, N, k, shape) in enumerate(BUILD):
        rng = np.random.default_rng(2000 + i)
        cfg = O.RbfConfig(kernel=kernel, polynomial_degree=deg, shape_parameter=shape)
        S = rng.random((N, n))
        V = np.stack([np.sum(S**2, 1), np.sum(np.sin(3 * S), 1)], 1)[:, :k]
        m = O.build_model(S, V, cfg)
        X = np.vstack((rng.random((9, n)), S[:2]))
        Y = np.array([m.eval(xx) for xx in X]); J = np.array([m.jac(xx) for xx in X])
        np.savez(os.path.join(OUT, name + ".npz"), kind="model", kernel=kernel, deg=deg, shape=shape, sites=S, values=V, X=X, Y=Y, J=J,
                 w=m.w, lam=m.lam, cond=m.cond)
    from oracle import descent_oracle as D
    for i, (name, n, k, boxed, normalize) in enumerate(DESCENT):          # steepest-descent LP (descent.jl:75-135), HiGHS
        rng = np.random.default_rng(3000 + i)
        B = 8
        jac = rng.normal(size=(B, k, n)) * rng.choice([1.0, 1e-3, 20.0], size=(B, 1, 1))
        x = rng.random((B, n))
        lb = np.full(n, 0.0 if boxed else -np.inf); ub = np.full(n, 1.0 if boxed else np.inf)
        if boxed:
            x[::3, 0] = 0.0; x[1::3, -1] = 1.0
        d = np.zeros((B, n)); om = np.zeros(B)
        for b in range(B):
            d[b], om[b] = D.lp_highs(x[b], jac[b], lb, ub, normalize)
        np.savez(os.path.join(OUT, name + ".npz"), kind="descent", jac=jac, x=x, lb=lb, ub=ub, normalize=normalize, d=d, omega=om)
    from oracle import iterate_oracle as IO
    for name, x0s, max_iter in TRAJ:                                        # whole optimize runs (algorithm.jl:919-958), two parabolas
        f = lambda z: np.array([np.sum((np.asarray(z) - 1.0) ** 2), np.sum((np.asarray(z) + 1.0) ** 2)])
        x0s = np.asarray(x0s, np.float64); n = x0s.shape[1]
        T = max_iter + 1
        rec = dict(ret=np.full((len(x0s), T), -1), it_stat=np.full((len(x0s), T), -1), x_index=np.zeros((len(x0s), T), int),
                   n_db=np.zeros((len(x0s), T), int), delta=np.zeros((len(x0s), T)), x=np.zeros((len(x0s), T, n)), fx=np.zeros((len(x0s), T, 2)),
                   rho=np.full((len(x0s), T), np.nan), omega=np.full((len(x0s), T), np.nan), n_train=np.zeros((len(x0s), T), int),
                   knife=np.zeros((len(x0s), T), bool), wall_tie=np.zeros((len(x0s), T), bool), n_iter=np.zeros(len(x0s), int))
        for b, x0 in enumerate(x0s):
            run = IO.optimize(f, x0, np.full(n, -np.inf), np.full(n, np.inf), O.RbfConfig(kernel="cubic"), IO.AlgoConfig(max_iter=max_iter))
            rec["n_iter"][b] = len(run.records)
            for t, r in enumerate(run.records):
                rec["ret"][b, t], rec["it_stat"][b, t], rec["x_index"][b, t], rec["n_db"][b, t] = r.ret_code, r.it_stat, r.x_index, r.n_db
                rec["delta"][b, t], rec["x"][b, t], rec["fx"][b, t], rec["rho"][b, t], rec["omega"][b, t] = r.delta, r.x, r.fx, r.rho, r.omega
                rec["n_train"][b, t], rec["knife"][b, t], rec["wall_tie"][b, t] = len(r.training_ids), r.knife, r.wall_tie
        np.savez(os.path.join(OUT, name + ".npz"), kind="trajectory", x0=x0s, max_iter=max_iter, **rec)
    print("wrote", len(SELECT) + len(BUILD) + len(DESCENT) + len(TRAJ), "fixtures to", OUT)


if __name__ == "__main__":
    main()
