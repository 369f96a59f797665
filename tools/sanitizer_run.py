"""Drives every kernel family of libmorbit_rbf.so once through the C ABI on small shapes -- the workload for
`compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck` (tools/sanitize.sh, logs under profiles/).

Small batches keep a racecheck pass (two orders of magnitude slower than a plain run) within minutes while still reaching:
rounds 1-3 (shared-memory and global-workspace variants), the three round-4 kernels (register-tiled elimination with its split-phase
mbarriers, blocked left-looking, literal), the four build kernels, every evaluation kernel (DMMA + cp.async.bulk ring, tiles, small,
wide, generic), Armijo batch, LP simplex, database append, model scatter, and a short lock-step run.
Results are checked against the oracle so that a sanitizer pass is also a correctness pass."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch                                        # noqa: E402
import morbit_jl_b200 as mb                          # noqa: E402
from morbit_jl_b200 import synthetic                 # noqa: E402
from morbit_jl_b200 import lockstep as LS            # noqa: E402
from morbit_jl_b200.multistart import MultistartBuilder, upload_batch   # noqa: E402
from oracle import c_oracle as CO                    # noqa: E402


def same_ids(res, ref, B):
    for b in range(B):
        for nm, cnt in (("r1", "n_r1"), ("r2", "n_r2"), ("r4", "n_r4")):
            assert list(getattr(res, nm)[b, :getattr(res, cnt)[b]]) == list(getattr(ref, nm)[b, :getattr(ref, cnt)[b]]), (b, nm)
        assert res.n_r3[b] == ref.n_r3[b]


def main():
    eng = mb.Engine(0)
    done = []
    # ---- rounds 1-4: schur path (<= 128 sites), block path (> 128 sites), literal path (budget-limited), global workspaces (n = 70)
    for (n, n_db, kernel, max_new, B) in [(30, 128, "multiquadric", 2**31 - 1, 4), (10, 200, "cubic", 2**31 - 1, 3),
                                          (6, 40, "gaussian", 1, 4), (70, 100, "cubic", 2**31 - 1, 1)]:
        cfg = mb.RbfConfig(kernel=kernel, max_model_points=-1 if n < 70 else 141)
        host = synthetic.multistart_batch(B, n=n, n_db=n_db, delta=0.1, func=synthetic.zdt3, local_fraction=0.4)
        host["max_new"][:] = max_new
        ref = CO.select_points_batched(cfg, host["sites"], host["x_index"], host["x"], host["delta"], host["delta_max"], host["glb"], host["gub"],
                                       False, False, host["max_new"], nthreads=4)
        res = eng.select_points(cfg, host["sites"], host["n_db"], host["x_index"], host["x"], host["delta"], host["delta_max"], host["glb"],
                                host["gub"], False, False, host["max_new"])
        same_ids(res, ref, B)
        # kept factorisation -> build (schur / prepared / general routes) -> evaluation
        dev = upload_batch(host, "cuda:0")
        builder = MultistartBuilder(eng, cfg, host["delta_max"])
        model, sel, status = builder.step(dev)
        model, sel, status = builder.step(dev, recycle=model)
        eng.sync()
        if max_new > 1:
            assert int((status != 0).sum().item()) == 0
        X = host["x"][:, None, :] + 0.05 * (np.random.default_rng(0).random((B, 40, n)) - 0.5)
        Y, J = eng.eval(model, X, True, True)            # DMMA Jacobian (n <= 64) / generic (n > 64)
        Y2, _ = eng.eval(model, X, True, False)          # DMMA values / wide
        Y3, J3 = eng.eval(model, X[:, :3], True, True)   # eval_small
        assert np.all(np.isfinite(Y)) and np.abs(Y - Y2).max() <= 1e-9 * max(1.0, np.abs(Y).max())
        assert np.abs(Y3 - Y[:, :3]).max() <= 1e-9 * max(1.0, np.abs(Y).max())
        # Armijo batch + LP
        d, omega, it, st = eng.descent_direction(J[:, 0], host["x"], host["glb"], host["gub"], True)
        dn = np.abs(d).max(axis=1, keepdims=True); dn[dn == 0] = 1.0
        eng.backtrack(model, host["x"], d / dn, dn[:, 0], omega)
        model.free()
        done.append(f"select/build/eval n={n} n_db={n_db} {kernel}")
    # ---- from-scratch build kernel (shared-memory and global-workspace systems) + tile kernels without the tensor-path copy
    rng = np.random.default_rng(1)
    for (n, N, k, kernel) in [(5, 21, 2, "cubic"), (30, 128, 2, "multiquadric"), (30, 200, 2, "cubic"), (100, 150, 2, "multiquadric")]:
        cfg = mb.RbfConfig(kernel=kernel)
        S = rng.random((2, N, n)); V = np.stack([np.sum(S**2, -1), np.sum(np.sin(S), -1)], -1)[..., :k]
        wr, lr, st = CO.build_batched(cfg, S, V, [N, N - 3])
        model, status = eng.build(cfg, S, V, [N, N - 3])
        X = rng.random((2, 33, n))
        Y, J = eng.eval(model, X, True, True)
        for b, Nb in enumerate((N, N - 3)):
            Yr = CO.eval_points(cfg, S[b, :Nb], wr[b, :Nb], lr[b], X[b])
            assert np.abs(Y[b] - Yr).max() <= 1e-8 * np.abs(Yr).max(), (n, N, np.abs(Y[b] - Yr).max())
        model.free()
        done.append(f"build/eval n={n} N={N}")
    # ---- device-resident database + lock-step optimize (db_append, model_scatter, sub-batches, criticality loop)
    x0 = np.random.default_rng(2).uniform(-3, 3, (6, 2))
    drv = LS.LockstepDriver(mb.RbfConfig(kernel="cubic"), synthetic.two_parabolas, x0, np.full(2, -np.inf), np.full(2, np.inf),
                            LS.AlgorithmConfig(max_iter=4))
    drv.run()
    x0 = synthetic.halton(4, 8)
    drv = LS.LockstepDriver(mb.RbfConfig(kernel="multiquadric"), synthetic.zdt3, x0, np.zeros(8), np.ones(8), LS.AlgorithmConfig(max_iter=3),
                            capacity=96)
    drv.run()
    torch.cuda.synchronize()
    done.append("lockstep")
    print("sanitizer workload ok:", "; ".join(done), f"| launches={eng.launch_count}")


if __name__ == "__main__":
    main()
