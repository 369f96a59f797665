"""From-scratch coefficient solve (mrbf_build_dev = update_model -> RBF.RBFInterpolationModel, RbfModel.jl:743-767) for a batch of
gathered training sets: systems/s and fraction of the FP64 peak on the reference-equivalent flop count.
    python tools/build_bench.py > profiles/build_bench_r01.json"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def flops(N, n, k):
    return 0.5 * N * N * (3 * n + 1) + (2.0 / 3.0) * (N + n + 1) ** 3 + 2 * k * (N + n + 1) ** 2


def main():
    import torch
    import morbit_jl_b200 as mb
    stream = torch.cuda.Stream()
    eng = mb.Engine(0, stream=stream.cuda_stream)
    rows = []
    configs = ((4096, 30, 61, 2, "multiquadric"), (4096, 30, 128, 2, "multiquadric"), (4096, 10, 66, 2, "cubic"), (1024, 30, 256, 2, "multiquadric"))
    if len(sys.argv) > 1:
        configs = (configs[int(sys.argv[1])],)
    for (B, n, N, k, kern) in configs:
        rng = np.random.default_rng(N)
        cfg = mb.RbfConfig(kernel=kern)
        x = 0.3 + 0.4 * rng.random((B, 1, n))
        S = np.clip(x + 0.3 * (rng.random((B, N, n)) - 0.5), 0, 1); S[:, 0] = x[:, 0]
        V = np.stack([np.sum(S ** 2, -1), np.sum(np.sin(3 * S), -1)], -1)[..., :k]
        dS, dV = torch.from_numpy(S).cuda(), torch.from_numpy(np.ascontiguousarray(V)).cuda()
        dN = torch.full((B,), N, dtype=torch.int32, device="cuda")
        with torch.cuda.stream(stream):
            model = None
            for _ in range(2):
                model, status = eng.build_dev(cfg, dS, dV, dN, None, None, recycle=model)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record(stream)
            for _ in range(reps):
                model, status = eng.build_dev(cfg, dS, dV, dN, None, status, recycle=model)
            e1.record(stream); stream.synchronize()
        ms = e0.elapsed_time(e1) / reps
        rows.append({"B": B, "n": n, "N": N, "k": k, "kernel": kern, "ms": ms, "systems_per_s": B / (ms * 1e-3), "ok": int((status == 0).sum().item()),
                     "tflops_reference_equivalent": B * flops(N, n, k) / (ms * 1e-3) / 1e12})
        model.free()
    print(json.dumps({"what": "from-scratch batched build (null-space method, one CTA per system)", "nt_env": os.environ.get("MRBF_BUILD_NT"), "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
