"""Runs the lock-step driver on the C3 shape and prints stop codes / timing (diagnostics)."""
import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
import torch
import morbit_jl_b200 as mb
from morbit_jl_b200 import lockstep as L, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 30
n = 30
x0 = synthetic.halton(B, n)
t0 = time.time()
drv = L.LockstepDriver(mb.RbfConfig(kernel="multiquadric"), synthetic.zdt3, x0, np.zeros(n), np.ones(n), L.AlgorithmConfig(max_iter=maxit), capacity=128, record=True)
torch.cuda.synchronize(); t1 = time.time()
x, fx, ret = drv.run()
torch.cuda.synchronize(); t2 = time.time()
print("init %.3f s, run %.3f s, lock-step iterations %d" % (t1 - t0, t2 - t1, drv.iter_counter - 1))
print("ret codes", dict(zip(*np.unique(ret, return_counts=True))))
print("iters done mean", drv.iters_done.double().mean().item(), "evals mean", drv.num_evals.double().mean().item(), "n_db max", drv.n_db.max().item())
print("numeric log", drv.numeric_log[:6], "build failures: instances", int((drv.build_failures > 0).sum()), "events", int(drv.build_failures.sum()))
print("func calls", drv.n_func_calls, "sites evaluated", drv.n_sites_evaluated, "launches", drv.engine.launch_count)
f0 = synthetic.zdt3(x0)
print("mean f0", f0.mean(0), "mean fx", fx.mean(0), "monotone", bool(np.all(fx <= f0 + 1e-12)))
if len(sys.argv) > 3:
    # debug: rerun, capturing the training sets of failing from-scratch builds in _improve
    drv = L.LockstepDriver(mb.RbfConfig(kernel="multiquadric"), synthetic.zdt3, x0, np.zeros(n), np.ones(n), L.AlgorithmConfig(max_iter=maxit), capacity=128, record=True)
    E = drv.engine
    orig = E.build_dev
    seen = [0]
    def hook(cfg, sites, values, N, shape=None, status=None, recycle=None):
        m, st = orig(cfg, sites, values, N, shape, status, recycle)
        bad = (st != 0).nonzero().flatten()
        if bad.numel() and seen[0] < 2 and sites.shape[1] > 40:
            seen[0] += 1
            b = int(bad[0]); Nb = int(N[b]); S_ = sites[b, :Nb].cpu().numpy(); V_ = values[b, :Nb].cpu().numpy()
            D = np.abs(S_[:, None, :] - S_[None, :, :]).max(-1) + np.eye(Nb) * 9
            i, j = np.unravel_index(np.argmin(D), D.shape)
            print("improve build failed: iter", drv.iter_counter, "N", Nb, "code", int(st[b]), "min pair dist", D.min(), "pair", i, j, "nan values", int(np.isnan(V_).sum()),
                  "outside box", int(((S_ < 0) | (S_ > 1)).any(1).sum()))
            np.set_printoptions(linewidth=250, precision=3)
            print("  centre", S_[0]); print("  site i - centre", S_[i] - S_[0]); print("  site j - centre", S_[j] - S_[0])
            # which instance is it?
            cand = (drv.x == sites[b, 0]).all(1).nonzero().flatten()
            bb = int(cand[0]); print("  instance", bb, "n_dirs", int(drv.n_dirs[bb]), "n_r1", int(drv.n_r1[bb]), "delta", float(drv.delta[bb]), "dirs[0]", drv.dirs[bb, 0].cpu().numpy(), "dirs[1]", drv.dirs[bb, 1].cpu().numpy())
        return m, st
    E.build_dev = hook
    drv.run()
