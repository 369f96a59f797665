"""BASELINE config C5: surrogate eval / Jacobian sweep, M = 1e3..1e7 trial points x 512 centres, d = 50, Gaussian and cubic,
k in {1, 2}.  Prints one JSON object (device-resident inputs, CUDA events on the engine's stream; M * d * 8 B > L2 from 1e6 on,
smaller M are timed with an L2 flush between launches).
    python tools/eval_sweep.py > profiles/eval_sweep_r01.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import morbit_jl_b200 as mb
    from morbit_jl_b200 import synthetic
    stream = torch.cuda.Stream()
    eng = mb.Engine(0, stream=stream.cuda_stream)
    peak = json.load(open(os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")))["peak_used_tflops"]
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")
    rows = []
    for kernel in ("gaussian", "cubic"):
        for k in (1, 2):
            centers, vals, _ = synthetic.eval_sweep(512, 50, k, 8, seed=0)
            cfg = mb.RbfConfig(kernel=kernel, shape_parameter=1.0 if kernel == "gaussian" else float("nan"))
            model, _ = eng.build(cfg, centers[None], vals[None], [512])
            cbar = torch.from_numpy(centers.mean(0)).cuda()
            lo, hi = torch.clamp(cbar - 0.2, min=0.0), torch.clamp(cbar + 0.2, max=1.0)
            for M in (10**3, 10**4, 10**5, 10**6, 10**7):
                g = torch.Generator(device="cuda"); g.manual_seed(1)
                X = (lo + (hi - lo) * torch.rand((1, M, 50), dtype=torch.float64, device="cuda", generator=g)).contiguous()
                Y = torch.empty((1, M, k), dtype=torch.float64, device="cuda")
                for want_j in (False, True):
                    if want_j and M * k * 50 * 8 > 8e9:
                        continue
                    J = torch.empty((1, M, k, 50), dtype=torch.float64, device="cuda") if want_j else None
                    reps = 20 if M <= 10**5 else (5 if M <= 10**6 else 2)
                    times = []
                    with torch.cuda.stream(stream):
                        for _ in range(3):
                            eng.eval_dev(model, X, Y, J)
                        for _ in range(reps):
                            if M * 50 * 8 < 200e6:
                                flush.zero_()
                            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            e0.record(stream); eng.eval_dev(model, X, Y, J); e1.record(stream)
                            stream.synchronize()
                            times.append(e0.elapsed_time(e1))
                    ms = float(np.median(times))
                    flop = 512 * (3 * 50 + 1 + (k * (1 + 2 * 50) if want_j else 2 * k)) + (k * 50 if want_j else 2 * 51 * k)
                    rows.append({"kernel": kernel, "k": k, "M": M, "what": "values+jacobian" if want_j else "values", "ms": ms,
                                 "points_per_s": M / (ms * 1e-3), "tflops": M * flop / (ms * 1e-3) / 1e12,
                                 "frac_fp64_peak": M * flop / (ms * 1e-3) / 1e12 / peak})
                    del J
                del X, Y
            model.free()
    print(json.dumps({"config": "C5: M trial points x 512 centres, d = 50", "fp64_peak_tflops": peak, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
