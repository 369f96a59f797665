"""Phase breakdown of the lock-step multistart run (C3 shape): LockstepDriver(profile=True) synchronises around every phase."""
import sys, time, json, numpy as np
sys.path.insert(0, "/root/repo")
import torch
import morbit_jl_b200 as mb
from morbit_jl_b200 import lockstep as L, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 20
n = 30
x0 = synthetic.halton(B, n)
out = {}
eng = mb.Engine(0, stream=torch.cuda.current_stream().cuda_stream)
scratch = {}
for prof in (None, False, True):          # None: warm-up run (allocates the per-size scratch buffers)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    drv = L.LockstepDriver(mb.RbfConfig(kernel="multiquadric"), synthetic.zdt3, x0, np.zeros(n), np.ones(n), L.AlgorithmConfig(max_iter=maxit), capacity=128, profile=bool(prof),
                           engine=eng, scratch=scratch)
    drv.run()
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    if prof:
        out["profiled_wall_s"] = t
        out["phases_s"] = {k: round(v, 4) for k, v in sorted(drv.phase_s.items(), key=lambda kv: -kv[1])}
        out["other_s"] = round(t - sum(drv.phase_s.values()), 4)
    elif prof is None:
        out["cold_wall_s"] = t
    else:
        out["wall_s"] = t
    out["instance_iterations"] = int(drv.iters_done.sum()); out["lockstep_iterations"] = drv.iter_counter - 1
    out["launches"] = drv.engine.launch_count; out["func_calls"] = drv.n_func_calls; out["sites_evaluated"] = drv.n_sites_evaluated
print(json.dumps(out, indent=1))
