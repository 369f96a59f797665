"""The heterogeneous batch of bench.py (mixed radii, round-3 budgets {0, 2, unlimited} = {5, 5, 90} %) on its own: per-kernel times and
how many instances enter round 4 under-poised (N0 < p) and by how much -- the workload of the literal kernel's prefix run.
    python tools/hetero_probe.py [B]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import morbit_jl_b200 as mb
    from morbit_jl_b200 import synthetic
    from morbit_jl_b200.multistart import MultistartBuilder, upload_batch
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    n = 30
    stream = torch.cuda.Stream()
    eng = mb.Engine(0, stream=stream.cuda_stream)
    cfg = mb.RbfConfig(kernel="multiquadric")
    host = synthetic.multistart_batch(B, n=n, n_db=128, delta=0.1, delta_max=0.5, func=synthetic.zdt3, local_fraction=0.5)
    rng = np.random.default_rng(12345)
    host["delta"] = 0.1 * rng.choice([0.25, 0.5, 1.0, 2.0], size=B)
    host["max_new"] = rng.choice([0, 2, 2**31 - 1], size=B, p=[0.05, 0.05, 0.9]).astype(np.int32)
    host["flags_in"][:, 0] = rng.integers(0, 2, size=B)
    dev = upload_batch(host, "cuda:0")
    bld = MultistartBuilder(eng, cfg, 0.5)
    with torch.cuda.stream(stream):
        m = None
        for _ in range(2):
            m, sel, st = bld.step(dev, recycle=m); stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            m, sel, st = bld.step(dev, recycle=m)
        e1.record(stream); stream.synchronize()
        eng.profile_enable(True)
        m, sel, st = bld.step(dev, recycle=m)
        prof = eng.profile_read(); eng.profile_enable(False)
    N0 = (1 + sel.n_r1 + sel.n_r2 + sel.n_r3).cpu().numpy()
    under = N0 < n + 1
    need = (n + 1 - N0)[under]
    print(json.dumps({"B": B, "ms_per_step": e0.elapsed_time(e1) / 5, "kernel_ms": {k: round(v, 3) for k, v in prof.items() if k != "eval"},
                      "under_poised_instances": int(under.sum()), "points_to_poised": {"mean": float(need.mean()) if need.size else 0.0,
                                                                                      "max": int(need.max()) if need.size else 0},
                      "builds_ok": int((st == 0).sum().item())}))


if __name__ == "__main__":
    main()
