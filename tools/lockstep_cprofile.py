"""Host-side profile (cProfile) of one warm lock-step run: where the wall time outside the kernels goes."""
import sys, cProfile, pstats, io, numpy as np
sys.path.insert(0, "/root/repo")
import torch
import morbit_jl_b200 as mb
from morbit_jl_b200 import lockstep as L, synthetic
B, n, maxit = 4096, 30, 20
x0 = synthetic.halton(B, n)
eng = mb.Engine(0, stream=torch.cuda.current_stream().cuda_stream)
scratch = {}
mk = lambda: L.LockstepDriver(mb.RbfConfig(kernel="multiquadric"), synthetic.zdt3, x0, np.zeros(n), np.ones(n), L.AlgorithmConfig(max_iter=maxit), capacity=128, engine=eng, scratch=scratch)
mk().run()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
drv = mk(); drv.run(); torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:4500])
