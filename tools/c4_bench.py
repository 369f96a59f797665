"""BASELINE config C4 (shape): n = 200 convex quadratics, 5 RBF outputs in one group, cubic kernel (default exponent 3 and
`shape_parameter = 1.0` => exponent 1, RbfModel.jl:683-684), max_model_points = 2n+1 = 401, theta_enlarge_1 = 2, theta_pivot = 1/4
(examples/large_scale_benchmarks.jl:154-160).  Device-resident batches of B database snapshots (400 sites each):
rounds 1-4 + coefficient solve per instance, CUDA events; the C port of the reference path on one host core beside it.
    python tools/c4_bench.py > profiles/c4_r01.json"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def batch(B, n, k, n_db, seed=0):
    rng = np.random.default_rng(seed)
    a = rng.random((k, n)); D = 0.5 + rng.random((k, n))
    func = lambda X: np.stack([np.sum(D[j] * (X - a[j]) ** 2, axis=-1) for j in range(k)], axis=-1)
    x = 0.3 + 0.4 * rng.random((B, n))
    sites = np.zeros((B, n_db, n))
    sites[:, 0] = x
    r = rng.random((B, n_db - 1, 1)) ** 0.25          # radii spread over box 2 (theta_2 * delta_max = 1), denser near the wall
    sites[:, 1:] = np.clip(x[:, None, :] + (rng.random((B, n_db - 1, n)) * 2 - 1) * 0.4 * r, 0.0, 1.0)
    return dict(sites=sites, values=func(sites), n_db=np.full(B, n_db, np.int32), x_index=np.ones(B, np.int32), x=x, delta=np.full(B, 0.1),
                glb=np.zeros(n), gub=np.ones(n), flags_in=np.tile(np.array([[1, 0]], np.int32), (B, 1)), max_new=np.full(B, 2**31 - 1, np.int32)), func


def main():
    import torch
    import morbit_jl_b200 as mb
    from morbit_jl_b200.multistart import MultistartBuilder, upload_batch
    from oracle import c_oracle as CO
    n, k, n_db = 200, 5, 400
    stream = torch.cuda.Stream()
    eng = mb.Engine(0, stream=stream.cuda_stream)
    rows = []
    for shape in ((float("nan"), 1.0) if len(sys.argv) < 2 else (float("nan"),)):
        cfg = mb.RbfConfig(kernel="cubic", shape_parameter=shape, max_model_points=2 * n + 1, theta_enlarge_1=2.0, theta_pivot=0.25)
        for B in ((1, 64, 592) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1].split(','))):
            host, func = batch(B, n, k, n_db)
            dev = upload_batch(host)
            builder = MultistartBuilder(eng, cfg, 0.5)
            with torch.cuda.stream(stream):
                model = None
                for _ in range(2):
                    model, sel, status = builder.step(dev, recycle=model)
                stream.synchronize()
                eng.profile_enable(True)
                model, sel, status = builder.step(dev, recycle=model)
                prof = eng.profile_read()
                eng.profile_enable(False)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 5
                e0.record(stream)
                for _ in range(reps):
                    model, sel, status = builder.step(dev, recycle=model)
                e1.record(stream); stream.synchronize()
                ms = e0.elapsed_time(e1) / reps
                # descent step on the fitted models: Jacobian at the iterate, LP, Armijo batch
                Jc = torch.empty((B, 1, k, n), dtype=torch.float64, device="cuda")
                d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                eng.eval_dev(model, dev.x.reshape(B, 1, n).contiguous(), None, Jc)
                out = eng.descent_direction_dev(Jc.view(B, k, n), dev.x, dev.glb, dev.gub, True)
                dn = out[0].abs().amax(dim=1).clamp_min(1e-300); dirn = (out[0] / dn[:, None]).contiguous()
                bt = eng.backtrack_dev(model, dev.x, dirn, dn, out[1])
                d0.record(stream)
                for _ in range(reps):
                    eng.eval_dev(model, dev.x.reshape(B, 1, n).contiguous(), None, Jc)
                    out = eng.descent_direction_dev(Jc.view(B, k, n), dev.x, dev.glb, dev.gub, True, out)
                    bt = eng.backtrack_dev(model, dev.x, dirn, dn, out[1], out=bt)
                d1.record(stream); stream.synchronize()
                ms_desc = d0.elapsed_time(d1) / reps
            Ntr = (1 + sel.n_r1 + sel.n_r2 + sel.n_r3 + sel.n_r4).double().mean().item()
            row = {"exponent": 3 if shape != shape else 1, "B": B, "n": n, "k": k, "db_sites": n_db, "max_model_points": 2 * n + 1,
                   "mean_training_points": Ntr, "builds_ok": int((status == 0).sum().item()), "ms_per_step": ms, "builds_per_s": B / (ms * 1e-3),
                   "kernel_ms": {k_: round(v, 3) for k_, v in prof.items() if v > 0}, "descent_step_ms": ms_desc,
                   "lp_ok": int((out[3] == 0).sum().item())}
            if B == 64:
                t0 = time.perf_counter()
                ns = 4
                ref = CO.select_points_batched(cfg, host["sites"][:ns], host["x_index"][:ns], host["x"][:ns], host["delta"][:ns], 0.5, host["glb"], host["gub"],
                                               True, False, 2**31 - 1, nthreads=1)
                t_sel = (time.perf_counter() - t0) / ns
                same = all(list(ref.r4[b, :ref.n_r4[b]]) == list(sel.r4[b, :sel.n_r4[b]].cpu().numpy()) and list(ref.r1[b, :ref.n_r1[b]]) == list(sel.r1[b, :sel.n_r1[b]].cpu().numpy())
                           for b in range(ns))
                row["cpu_port_1core_select_ms"] = t_sel * 1e3
                row["indices_equal_to_c_port_first4"] = bool(same)
            rows.append(row)
            model.free()
            del dev, builder
            torch.cuda.empty_cache()
    print(json.dumps({"what": "C4 shape (n = 200, 5 outputs, cubic, 401 model points, 400-site snapshots): rounds 1-4 + solve, device-resident batches",
                      "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
