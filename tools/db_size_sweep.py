"""Builds/s of the C3 batch (4096 ZDT3 instances, n = 30, k = 2, multiquadric) for database snapshots of 31 / 128 / 512 sites
(SURVEY 8(d)): device-resident inputs, CUDA events on the engine's stream, per-kernel times from the instrumented pass.
    python tools/db_size_sweep.py > profiles/db_size_sweep_r01.json"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import morbit_jl_b200 as mb
    from morbit_jl_b200 import synthetic
    from morbit_jl_b200.multistart import MultistartBuilder, upload_batch
    stream = torch.cuda.Stream()
    eng = mb.Engine(0, stream=stream.cuda_stream)
    cfg = mb.RbfConfig(kernel="multiquadric")
    rows = []
    sizes = (31, 128, 512) if len(sys.argv) < 2 else (int(sys.argv[1]),)
    for n_db in sizes:
        B = 4096 if len(sys.argv) < 3 else int(sys.argv[2])
        host = synthetic.multistart_batch(B, n=30, n_db=n_db, delta=0.1, delta_max=0.5, func=synthetic.zdt3)
        dev = upload_batch(host, "cuda:0")
        builder = MultistartBuilder(eng, cfg, 0.5, func=synthetic.zdt3)          # new round-3 sites get their true values
        model = None
        with torch.cuda.stream(stream):
            for _ in range(2):
                model, sel, status = builder.step(dev, recycle=model)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            stream.synchronize(); e0.record(stream)
            for _ in range(3):
                model, sel, status = builder.step(dev, recycle=model)
            e1.record(stream); stream.synchronize()
        ms = e0.elapsed_time(e1) / 3
        eng.profile_enable(True)
        with torch.cuda.stream(stream):
            model, sel, status = builder.step(dev, recycle=model)
        prof = eng.profile_read(); eng.profile_enable(False)
        N = 1 + sel.n_r1 + sel.n_r2 + sel.n_r3 + sel.n_r4
        rows.append({"db_sites": n_db, "ms_per_step": ms, "builds_per_s": B / (ms * 1e-3), "mean_training_points": float(N.double().mean().item()),
                     "mean_new_round3_sites": float(sel.n_r3.double().mean().item()), "builds_ok": int((status == 0).sum().item()),
                     "kernel_ms": {k: round(v, 3) for k, v in prof.items() if k != "eval"}})
        model.free(); del builder, dev
        torch.cuda.empty_cache()
    print(json.dumps({"config": "C3: 4096 ZDT3 instances, n = 30, k = 2, multiquadric, default RbfConfig", "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
