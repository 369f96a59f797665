"""Run-to-run determinism under contention -- a stand-in for `compute-sanitizer --tool racecheck`, which this pool does not offer.

No kernel of the library uses floating-point atomics, so two runs of the same call must agree BIT FOR BIT whatever else the GPU is
doing; a difference is a data race (a missing barrier, a ring slot read after its reuse) made visible by a shifted schedule.  The
script repeats the hot path on fixed inputs while a second host thread keeps the device busy through its own context, and counts
differing outputs per component:

  A. host-buffer entry points on small ragged batches (the shapes of tests/test_gpu_plugin.py::test_reentrant_contexts_from_concurrent_threads):
     mrbf_select_points -> mrbf_build (reduced-system route and QR fallback) -> mrbf_model_coeffs / mrbf_eval
  B. the headline batch (C3: 4096 instances, n = 30, 128 sites) through the device entry points, kept-factorisation route and
     two-phase route, coefficients of all instances compared.

usage: python tools/determinism_stress.py [repetitions_A] [repetitions_B]      (prints one JSON line)
"""
import json
import sys
import threading

import numpy as np

sys.path.insert(0, ".")
import morbit_jl_b200 as mb                     # noqa: E402
from morbit_jl_b200 import synthetic            # noqa: E402
from morbit_jl_b200.multistart import MultistartBuilder, upload_batch   # noqa: E402


def small_case(eng, seed, rep, B=48):
    cfg = mb.RbfConfig(kernel="cubic" if seed % 2 else "multiquadric")
    h = synthetic.multistart_batch(B, n=8 + seed, n_db=40 + 7 * rep, delta=0.1, func=synthetic.zdt3, local_fraction=0.5,
                                   first_instance=100 * seed + rep)
    res = eng.select_points(cfg, h["sites"], h["n_db"], h["x_index"], h["x"], h["delta"], h["delta_max"], h["glb"], h["gub"])
    n = h["sites"].shape[2]
    N = 1 + res.n_r1 + res.n_r2 + res.n_r3 + res.n_r4
    ts = int(N.max())
    S = np.zeros((B, ts, n)); V = np.zeros((B, ts, 2))
    for b in range(B):
        ids = [int(h["x_index"][b])] + list(res.r1[b, :res.n_r1[b]]) + list(res.r2[b, :res.n_r2[b]])
        P = np.vstack([h["sites"][b, np.array(ids) - 1], res.r3_sites[b, :res.n_r3[b]].reshape(-1, n),
                       h["sites"][b, res.r4[b, :res.n_r4[b]].astype(int) - 1].reshape(-1, n)])
        S[b, :len(P)] = P; V[b, :len(P)] = synthetic.zdt3(P)
    model, status = eng.build(cfg, S, V, N, raise_on_failure=False)
    w, lam = model.coeffs()
    Y, J = eng.eval(model, h["x"][:, None, :], True, True)
    model.free()
    valid = lambda a, c: np.concatenate([a[b, :c[b]].ravel() for b in range(B)])
    return dict(r1=valid(res.r1, res.n_r1), r2=valid(res.r2, res.n_r2), r3=valid(res.r3_sites, res.n_r3), r4=valid(res.r4, res.n_r4),
                counts=np.stack([res.n_r1, res.n_r2, res.n_r3, res.n_r4]), status=status.copy(),
                w=valid(w, N), lam=lam.copy(), values=Y.copy(), jacobians=J.copy())


def differing(a, b):
    return [k for k in a if not (np.shape(a[k]) == np.shape(b[k]) and np.array_equal(a[k], b[k], equal_nan=True))]


def part_a(reps):
    cases = [(s, r) for s in (1, 2) for r in range(6)]
    eng = mb.Engine(0)
    ref = {c: small_case(eng, *c) for c in cases}
    eng.close()
    bad = {}
    lock = threading.Lock()

    def work(order):
        e = mb.Engine(0)
        for _ in range(reps):
            for c in order:
                d = differing(ref[c], small_case(e, *c))
                if d:
                    with lock:
                        for k in d:
                            bad[k] = bad.get(k, 0) + 1
        e.close()

    ths = [threading.Thread(target=work, args=(cases,)), threading.Thread(target=work, args=(cases[::-1],)),
           threading.Thread(target=work, args=(cases[3:] + cases[:3],))]
    for t in ths: t.start()
    for t in ths: t.join()
    return dict(calls=3 * reps * len(cases), instances_per_call=48, differing=bad)


def part_b(reps):
    import torch
    cfg = mb.RbfConfig(kernel="multiquadric")
    h = synthetic.multistart_batch(4096, n=30, n_db=128)
    stop = threading.Event()

    def disturb():                                  # another context keeps rounds 1-4 of a different batch in flight
        st_ = torch.cuda.Stream()
        torch.cuda.set_stream(st_)                  # (per-thread current stream)
        e = mb.Engine(0, stream=st_.cuda_stream)
        g = synthetic.multistart_batch(1024, n=20, n_db=96, first_instance=5000)
        d = upload_batch(g, "cuda:0")
        bld = MultistartBuilder(e, mb.RbfConfig(kernel="cubic"), g["delta_max"])
        m = None
        while not stop.is_set():
            m, _, _ = bld.step(d, fused=True, recycle=m)
            e.sync()
        e.close()

    eng = mb.Engine(0, stream=torch.cuda.current_stream().cuda_stream)
    d = upload_batch(h, "cuda:0")
    out = {}
    for fused in (True, False):
        bld = MultistartBuilder(eng, cfg, h["delta_max"])

        def run():
            m, sel, st = bld.step(d, fused=fused)
            eng.sync()
            w, lam = m.coeffs()
            torch.cuda.synchronize()
            head = lambda ids, cnt: torch.where(torch.arange(ids.shape[1], device=ids.device)[None, :] < cnt[:, None], ids,
                                                torch.zeros_like(ids)).cpu().numpy()      # entries beyond the counts are undefined
            r = dict(r2=head(sel.r2, sel.n_r2), r4=head(sel.r4, sel.n_r4),
                     counts=torch.stack([sel.n_r1, sel.n_r2, sel.n_r3, sel.n_r4]).cpu().numpy().copy(),
                     status=st.cpu().numpy().copy(), w=w, lam=lam)
            m.free()
            return r

        ref = run()
        th = threading.Thread(target=disturb); stop.clear(); th.start()
        bad = {}
        for _ in range(reps):
            for k in differing(ref, run()):
                bad[k] = bad.get(k, 0) + 1
        stop.set(); th.join()
        out["kept_factorisation" if fused else "two_phase"] = dict(calls=reps, instances_per_call=4096, differing=bad)
    eng.close()
    return out


if __name__ == "__main__":
    ra = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rb = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    print(json.dumps(dict(small_ragged_batches_three_threads=part_a(ra), headline_batch_with_disturber=part_b(rb))))
