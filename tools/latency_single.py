"""Single-instance (B = 1) latencies of the plugin calls -- the `optimize(mop, x0)` use case of BASELINE configs C1 / C2 --
next to the CPU oracle's C port on one core.  Host-pointer entry points (what the Julia shim calls), wall-clock medians.
    python tools/latency_single.py > profiles/latency_single_r01.json"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def med(f, reps=15):
    f(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts)) * 1e3


def main():
    import morbit_jl_b200 as mb
    from morbit_jl_b200 import synthetic
    from oracle import c_oracle as CO
    eng = mb.Engine(0)
    rows = []
    for n, n_db, mmp, kern in ((2, 12, -1, "cubic"), (30, 128, 61, "multiquadric"), (30, 128, -1, "multiquadric"), (30, 600, -1, "multiquadric")):
        cfg = mb.RbfConfig(kernel=kern, max_model_points=mmp)
        h = synthetic.multistart_batch(1, n=n, n_db=n_db, delta=0.1, func=synthetic.zdt3 if n > 2 else synthetic.two_parabolas)
        sel = lambda: eng.select_points(cfg, h["sites"], h["n_db"], h["x_index"], h["x"], h["delta"], h["delta_max"], h["glb"], h["gub"])
        res = sel()
        ids = [int(h["x_index"][0])] + list(res.r1[0, :res.n_r1[0]]) + list(res.r2[0, :res.n_r2[0]]) + list(res.r4[0, :res.n_r4[0]])
        P = np.vstack([h["sites"][0, np.array(ids) - 1], res.r3_sites[0, :res.n_r3[0]].reshape(-1, n)])
        V = (synthetic.zdt3 if n > 2 else synthetic.two_parabolas)(P)
        N = len(P)
        model = [None]
        def build():
            if model[0] is not None: model[0].free()
            model[0], _ = eng.build(cfg, P[None], V[None], [N])
        t_sel, t_build = med(sel), med(build)
        kept = [None, None]
        def sel_keep():
            kept[0], kept[1] = eng.select_points_keep(cfg, h["sites"], h["n_db"], h["x_index"], h["x"], h["delta"], h["delta_max"], h["glb"],
                                                      h["gub"], prepared=kept[1])
        sel_keep()
        r3v = (synthetic.zdt3 if n > 2 else synthetic.two_parabolas)(kept[0].r3_sites[0])[None]
        m2 = [None]
        def build_kept():
            m2[0], _ = eng.build_prepared(cfg, kept[1], h["sites"], h["values"], h["x_index"], kept[0], r3v, recycle=m2[0])
        t_selk, t_buildk = med(sel_keep), med(build_kept)
        m2[0].free(); kept[1].free()
        x = h["x"][:, None, :]
        t_eval = med(lambda: eng.eval(model[0], x, True, False), 50)
        t_jac = med(lambda: eng.eval(model[0], x, False, True), 50)
        d = np.ones((1, n)) / np.sqrt(n)
        t_bt = med(lambda: eng.backtrack(model[0], h["x"], d, 0.1, 0.5), 30)
        c_sel = med(lambda: CO.select_points_batched(cfg, h["sites"], h["x_index"], h["x"], h["delta"], h["delta_max"], h["glb"], h["gub"],
                                                      False, False, 2**31 - 1, nthreads=1), 5)
        c_build = med(lambda: CO.build_batched(cfg, P[None], V[None], [N], nthreads=1), 5)
        rows.append({"n": n, "db_sites": n_db, "max_model_points": mmp, "kernel": kern, "training_points": N,
                     "gpu_ms": {"select_points": t_sel, "build": t_build, "select_points_keep": t_selk, "build_prepared": t_buildk, "eval_1pt": t_eval, "jacobian_1pt": t_jac, "backtrack_118pts": t_bt},
                     "cpu_port_1core_ms": {"select_points": c_sel, "build": c_build}})
        model[0].free()
    print(json.dumps({"what": "B = 1 latencies through the host-pointer C ABI (wall clock, median)", "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
