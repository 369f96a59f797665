#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native RBF-surrogate hot path (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One step = one pass of the hot path over one batch: for every instance of BASELINE config C3
(4096 independent multistart ZDT3 n=30 instances per GPU, database snapshot of 128 sites, multiquadric
RbfConfig) run prepare_update_model (rounds 1-4) + training-set gather + update_model (batched build).
metric = RBF model builds/s, whole job.  Secondary: surrogate evals/s and Jacobians/s on config C5
(10^6 trial points x 512 centres, d = 50).  Weak scaling: every rank owns its own 4096 instances, no
collective on the data path; one final gather of per-instance result rows.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU = 4096
N_VARS, N_DB, K_OUT = 30, 128, 2
DELTA, DELTA_MAX = 0.1, 0.5
KERNEL = "multiquadric"
METRIC = "rbf_model_builds_per_s"
UNIT = "builds/s"
WORKLOAD = (f"C3: {B_PER_GPU} independent multistart ZDT3 n={N_VARS} k={K_OUT} instances per GPU, database snapshot "
            f"{N_DB} sites/instance, RbfConfig(kernel=:{KERNEL}) defaults; build = rounds 1-4 + coefficient solve "
            f"(from the kept round-4 factorisation)")


def load_synthetic():
    """morbit.jl_b200/synthetic.py is pure NumPy (seeded input generators shared by both arms).  The CPU arms load it by path so
    that the product package -- and with it libmorbit_rbf.so -- is never imported into the reference process."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_mrbf_synthetic", os.path.join(ROOT, "morbit.jl_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


LS_CODES = {0: "CONTINUE", 1: "MAX_ITER", 2: "BUDGET_EXHAUSTED", 3: "CRITICAL", 4: "TOLERANCE", 5: "INFEASIBLE", 6: "DB_FULL", 7: "NUMERIC"}


def fp64_peak_tflops():
    """Measured FP64 FMA-pipe peak of this pool's B200 (tools/fp64_peak.cu); MEASURED_PEAKS.json has no FP64 figure."""
    path = os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")
    try:
        d = json.load(open(path))
        return float(d["peak_used_tflops"]), "measured: tools/fp64_peak.cu (max of DFMA 33.8 / DMMA 37.2 TFLOP/s), profiles/fp64_peaks_r01.json"
    except Exception:
        return 37.0, "fallback: nominal B200 FP64 (no measured file)"


def build_flops(N, n, k):
    """SURVEY §8(d) primary figure: assembly + dense LU of the saddle system + solves."""
    N = np.asarray(N, dtype=np.float64)
    return 0.5 * N * N * (3 * n + 1) + (2.0 / 3.0) * (N + n + 1) ** 3 + 2 * k * (N + n + 1) ** 2


def round4_flops(N0, n_acc, n):
    """4 N^2 + 6 N (n+1) per candidate (four N-GEMV + sparse Givens), N growing with each acceptance."""
    tot = 0.0
    for a, c in zip(np.asarray(N0), np.asarray(n_acc)):
        Ns = a + np.arange(c, dtype=np.float64)
        tot += float(np.sum(4 * Ns * Ns + 6 * Ns * (n + 1)))
    return tot


def rounds123_flops(n_cand, n_picked, n):
    """4 n (n-j) per remaining candidate per filter step (two GEMV with Z in R^{n x (n-j)})."""
    tot = 0.0
    for c, p in zip(np.asarray(n_cand), np.asarray(n_picked)):
        for j in range(1, int(p) + 1):
            tot += 4.0 * n * (n - j) * max(c - j, 0)
    return tot


class ClockSampler:
    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def start(self):
        def run():
            q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            while not self._stop.is_set():
                try:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(",")])
                except Exception:
                    pass
                self._stop.wait(0.2)
        self._t = threading.Thread(target=run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def host_threads() -> int:
    """Host cores this process may use.  torchrun exports OMP_NUM_THREADS=1; the oracle's OpenMP loops take an explicit thread
    count, so the CPU arm still uses every core of the box (only rank 0 runs it)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_step(CO, cfg, host, n_threads, lo=0, hi=None):
    """One CPU step over instances [lo, hi): ONE call into the C port (oracle/rbf_oracle.c::orc_select_and_build_batched), which runs
    rounds 1-4, the training-set gather from the database arrays, the objective values of new round-3 sites and the saddle solve
    for one instance per thread -- no Python, no NumPy and nothing single-threaded between the two halves."""
    sl = slice(lo, hi)
    t0 = time.perf_counter()
    N, ids, w, lam, st = CO.select_and_build_batched(cfg, host["sites"][sl], host["values"][sl], host["x_index"][sl], host["x"][sl],
                                                     host["delta"][sl], host["delta_max"], host["glb"], host["gub"], False, False,
                                                     host["max_new"][sl], func="zdt3", nthreads=n_threads)
    return time.perf_counter() - t0, int((st == 0).sum())


def cpu_baseline(CO, cfg, host, n_threads, budget_s=15.0):
    """The oracle's C port on the host cores, one instance per thread, the SAME batch as the GPU arm, repeated for ~budget_s."""
    B = host["sites"].shape[0]
    cpu_step(CO, cfg, host, n_threads, 0, min(B, 4 * n_threads))          # warm-up (page-in, OpenMP team)
    reps, total, ok = 0, 0.0, 0
    while reps < 1 or (total < budget_s and reps < 50):
        t, ok = cpu_step(CO, cfg, host, n_threads)
        total += t; reps += 1
    return B * reps / total, B, reps, ok


def eval_sweep(eng, torch, stream, M, kernel, want_j, steps, warmup):
    """C5: M trial points x 512 centres, d = 50, k = 1; inputs resident in HBM (M*d*8 = 400 MB at 1e6 > L2)."""
    import morbit_jl_b200 as mb
    from morbit_jl_b200 import synthetic
    centers, vals, _ = synthetic.eval_sweep(512, 50, 1, 8, seed=0)
    cfg = mb.RbfConfig(kernel=kernel, shape_parameter=1.0 if kernel == "gaussian" else float("nan"))
    model, _ = eng.build(cfg, centers[None], vals[None], [512])
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    cbar = torch.from_numpy(centers.mean(0)).cuda()
    lo, hi = torch.clamp(cbar - 0.2, min=0.0), torch.clamp(cbar + 0.2, max=1.0)
    X = (lo + (hi - lo) * torch.rand((1, M, 50), dtype=torch.float64, device="cuda", generator=g)).contiguous()
    Y = torch.empty((1, M, 1), dtype=torch.float64, device="cuda")
    J = torch.empty((1, M, 1, 50), dtype=torch.float64, device="cuda") if want_j else None
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            eng.eval_dev(model, X, Y, J)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream.synchronize()
        e0.record(stream)
        for _ in range(steps):
            eng.eval_dev(model, X, Y, J)
        e1.record(stream)
        stream.synchronize()
    ms = e0.elapsed_time(e1) / steps
    model.free()
    return M / (ms * 1e-3), ms


def run_ours(args):
    import torch
    import torch.distributed as dist
    import morbit_jl_b200 as mb
    from morbit_jl_b200 import synthetic
    from morbit_jl_b200.multistart import MultistartBuilder, upload_batch, train_stride_for, gather_results

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    eng = mb.Engine(local, stream=stream.cuda_stream)
    cfg = mb.RbfConfig(kernel=KERNEL)
    B = args.instances
    host = synthetic.multistart_batch(B, n=N_VARS, n_db=N_DB, delta=DELTA, delta_max=DELTA_MAX, func=synthetic.zdt3,
                                      first_instance=rank * B)
    dev = upload_batch(host, f"cuda:{local}")
    builder = MultistartBuilder(eng, cfg, DELTA_MAX)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        model = None            # every step replaces the previous iteration's models in place (no allocation per step)
        for _ in range(args.warmup):
            model, sel, status = builder.step(dev, recycle=model); stream.synchronize()
        l0 = eng.launch_count
        sampler = ClockSampler(local); sampler.start()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            model, sel, status = builder.step(dev, recycle=model)
        e1.record(stream)
        stream.synchronize()
        barrier()
        launches = eng.launch_count - l0
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B / (ms * 1e-3)

    # ---- end to end through the public API with HOST buffers: pinned H2D of the step's inputs (the whole database snapshot,
    # as the Julia shim hands it over), D2H of the indices / flags / status; multistart.HostPipeline hides the copies of one
    # half of the batch behind the kernels of the other half
    from morbit_jl_b200.multistart import HostPipeline
    pipe = HostPipeline(eng, cfg, DELTA_MAX, host, f"cuda:{local}", stream, chunks=args.e2e_chunks, buffers=args.e2e_buffers, outputs=args.e2e_outputs)
    for _ in range(max(args.e2e_outputs * args.e2e_buffers, args.warmup // 2)):      # every buffer / output set used once (allocations)
        pipe.step(); pipe.drain(); stream.synchronize()
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(args.steps):
        _, e2e_outs = pipe.step()
    pipe.drain()                    # the last steps' result copies are inside the timed region
    t1.record(stream)
    stream.synchronize()
    barrier()
    e2e_ms = t0.elapsed_time(t1) / args.steps
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    e2e_status_ok = int(sum(int((o[-1].numpy() == 0).sum()) for o in e2e_outs))
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())

    def timed_pipeline(pipe_, nb):
        for _ in range(max(nb, args.warmup // 2)):          # every buffer / output set used once (allocations)
            pipe_.step(); pipe_.drain(); stream.synchronize()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(args.steps):
            _, outs_ = pipe_.step()
        pipe_.drain()
        a1.record(stream)
        stream.synchronize()
        barrier()
        ms_ = a0.elapsed_time(a1) / args.steps
        if world > 1:
            t_ = torch.tensor([ms_], dtype=torch.float64, device="cuda")
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            ms_ = float(t_.item())
        return ms_, int(sum(int((o[-1].numpy() == 0).sum()) for o in outs_))

    # ---- the same end-to-end step with the databases resident on the device: only iterate, radius, flags and
    # budget travel per step -- what changes between two model updates on one database (criticality loop, algorithm.jl:523-612)
    del pipe
    pipe_r = HostPipeline(eng, cfg, DELTA_MAX, host, f"cuda:{local}", stream, chunks=1, buffers=2, outputs=2, resident_db=True)
    e2e_res_ms, e2e_res_ok = timed_pipeline(pipe_r, 4)
    e2e_resident = {"value": world * B / (e2e_res_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_res_ms, "builds_ok": e2e_res_ok,
                    "h2d_bytes_per_step": int(pipe_r.h2d_bytes), "d2h_bytes_per_step": int(pipe_r.d2h_bytes),
                    "path": "databases uploaded once and kept on the device (mrbf_db_append_dev grows them in a driver); per step: pinned H2D of "
                            "iterate / radius / flags / budget -> mrbf_select_points_keep_dev + mrbf_build_prepared_dev -> D2H of indices / flags / status"}
    del pipe_r

    # ---- the headline end-to-end figure: databases resident on the device AND growing the way they do inside optimize -- every step
    # uploads one newly evaluated site + its values per instance (Databases.jl:390-401: every trial point is appended to every database)
    # and appends it with mrbf_db_append_dev, next to iterate / radius / flags / budget; results leave as before
    pipe_a = HostPipeline(eng, cfg, DELTA_MAX, host, f"cuda:{local}", stream, chunks=1, buffers=2, outputs=2, resident_db=True, append_rows=1)
    e2e_app_ms, e2e_app_ok = timed_pipeline(pipe_a, 4)
    e2e_append = {"value": world * B / (e2e_app_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_app_ms, "builds_ok": e2e_app_ok,
                  "h2d_bytes_per_step": int(pipe_a.h2d_bytes), "d2h_bytes_per_step": int(pipe_a.d2h_bytes), "appended_rows_per_instance": 1,
                  "path": "databases resident on the device; per step: pinned H2D of one new evaluated site + values per instance and of iterate / "
                          "radius / flags / budget / database size -> mrbf_db_append_dev -> mrbf_select_points_keep_dev + mrbf_build_prepared_dev -> "
                          "D2H of indices / flags / status (multistart.HostPipeline: uploads and result copies overlap the kernels of the "
                          "neighbouring steps; the timed region ends when the last copy has landed)"}
    del pipe_a

    # ---- strong scaling (the 4096 instances of config C3 split [g B / G, (g + 1) B / G) over the ranks)
    strong = None
    if world > 1:
        lo_s, hi_s = rank * B // world, (rank + 1) * B // world
        host_s = {k_: (v_[lo_s:hi_s] if isinstance(v_, np.ndarray) and v_.ndim >= 1 and v_.shape[0] == B else v_) for k_, v_ in
                  synthetic.multistart_batch(B, n=N_VARS, n_db=N_DB, delta=DELTA, delta_max=DELTA_MAX, func=synthetic.zdt3).items()}
        dev_s = upload_batch(host_s, f"cuda:{local}")
        b_s = MultistartBuilder(eng, cfg, DELTA_MAX)
        with torch.cuda.stream(stream):
            m_s = None
            for _ in range(args.warmup):
                m_s, _, _ = b_s.step(dev_s, recycle=m_s); stream.synchronize()
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(stream)
            for _ in range(args.steps):
                m_s, _, st_s = b_s.step(dev_s, recycle=m_s)
            s1.record(stream)
            stream.synchronize()
            barrier()
        ms_s = s0.elapsed_time(s1) / args.steps
        t_ = torch.tensor([ms_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        ms_s = float(t_.item())
        per = hi_s - lo_s
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        strong = {"metric": METRIC, "scaling": "strong", "value": B / (ms_s * 1e-3), "unit": UNIT, "ms_per_step": ms_s,
                  "instances_total": B, "instances_per_gpu": per,
                  "limiter": f"one CTA per instance: {per} CTAs on {sms} SMs = {per / (2 * sms):.2f} waves of the rounds-1-3 kernel (2 CTAs/SM), "
                             f"{per / (4 * sms):.2f} of the round-4 pre-kernel (4/SM), {per / (2 * sms):.2f} of the elimination kernel (2/SM) -- the last, partly "
                             "filled wave of each kernel is the loss against ideal strong scaling; there is no communication on the data path",
                  "config": {"workload": f"C3 strong: {B} instances in total, [g B / G, (g + 1) B / G) on rank g"}}
        m_s.free(); del dev_s, b_s

    # ---- per-kernel device times (one extra, untimed-for-the-metric step with event brackets) and roofline
    eng.profile_enable(True)
    with torch.cuda.stream(stream):
        model, sel, status = builder.step(dev, recycle=model)
    prof = eng.profile_read()
    eng.profile_enable(False)
    n_r1, n_r2, n_r3, n_r4 = (getattr(sel, a).cpu().numpy() for a in ("n_r1", "n_r2", "n_r3", "n_r4"))
    N0 = 1 + n_r1 + n_r2 + n_r3
    Ntrain = N0 + n_r4
    ok = int((status.cpu().numpy() == 0).sum())
    peak, peak_src = fp64_peak_tflops()
    kflops = {"rounds123": rounds123_flops(np.full(B, N_DB - 1), n_r1 + n_r2, N_VARS), "round4": round4_flops(N0, n_r4, N_VARS),
              "build": float(np.sum(build_flops(Ntrain, N_VARS, K_OUT))), "build_prepared": float(np.sum(build_flops(Ntrain, N_VARS, K_OUT)))}
    dom = max(("rounds123", "round4", "build", "build_prepared"), key=lambda k: prof[k])
    ach = kflops[dom] / (prof[dom] * 1e-3) / 1e12
    step_total = (prof["rounds123"] + prof["round4"] + prof["round4_fallback"] + prof.get("round4_prefix", 0.0) + prof["gather"] + prof["build"]
                  + prof["build_prepared"])
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic_r02.json"))).get(dom) if B == B_PER_GPU else None
    except Exception:
        traffic = None
    roofline = {"bound": "tensor", "pipe": "fp64 (DFMA; B200 FP64 tensor peak equals the FMA-pipe peak)", "kernel": dom,
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel_ms": {k: round(v, 4) for k, v in prof.items() if k != "eval"},
                "kernel_share_of_step": round(prof[dom] / step_total, 4) if step_total > 0 else None,
                "algorithmic_flops_per_launch": kflops[dom]}

    # ---- final gather of per-instance results (the only collective)
    rows = np.stack([Ntrain, n_r4, status.cpu().numpy()], axis=1).astype(np.float64)
    allrows = gather_results(rows, world * B, rank, world)
    gather_c_abi = None
    if world > 1:
        # the same gather through the C export a Julia host would call (mrbf_comm_init + mrbf_gather: ncclAllGather on the library's
        # own communicator; the 128-byte NCCL id travels over the host channel, here torch.distributed's store)
        from morbit_jl_b200.multistart import gather_results_c_abi
        box = [mb.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm = mb.Comm(local, box[0], rank, world)
        t0g = time.perf_counter()
        allrows_c = gather_results_c_abi(comm, rows, world * B)
        gather_c_abi = {"ms": (time.perf_counter() - t0g) * 1e3, "equal_to_torch_distributed": bool(np.array_equal(allrows_c, allrows))}
        comm.close()
        if not gather_c_abi["equal_to_torch_distributed"]:
            raise SystemExit("mrbf_gather disagrees with torch.distributed.all_gather")

    # ---- secondary metric: surrogate evals/s and Jacobians/s (config C5)
    secondary = []
    if args.eval_points > 0:
        fe = 512 * (3 * 50 + 1 + 2 * 1) + 2 * 51 * 1
        fj = 512 * (3 * 50 + 1 + 1 * (1 + 2 * 50)) + 50
        for kern in ("gaussian", "cubic"):
            ev, ev_ms = eval_sweep(eng, torch, stream, args.eval_points, kern, False, args.steps, args.warmup)
            jv, jv_ms = eval_sweep(eng, torch, stream, args.eval_points, kern, True, args.steps, args.warmup)
            if world > 1:
                t = torch.tensor([ev, jv], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                ev, jv = (float(v) * world for v in t.tolist())
            secondary.append({"metric": "surrogate_evals_per_s", "kernel": kern, "value": ev, "unit": "evals/s", "ms_per_step": ev_ms,
                              "config": {"workload": f"C5: {args.eval_points} trial points x 512 centres, d=50, k=1, per GPU"},
                              "roofline": {"bound": "tensor", "pipe": "fp64", "achieved": ev / world * fe / 1e12, "peak": peak,
                                           "unit": "TFLOP/s", "frac": ev / world * fe / 1e12 / peak, "flop_per_point": fe}})
            secondary.append({"metric": "surrogate_jacobians_per_s", "kernel": kern, "value": jv, "unit": "jacobians/s", "ms_per_step": jv_ms,
                              "config": {"workload": f"C5: {args.eval_points} trial points x 512 centres, d=50, k=1, values + Jacobian, per GPU"},
                              "roofline": {"bound": "tensor", "pipe": "fp64", "achieved": jv / world * fj / 1e12, "peak": peak,
                                           "unit": "TFLOP/s", "frac": jv / world * fj / 1e12 / peak, "flop_per_point": fj}})

    # ---- secondary metric: one steepest-descent step per instance on the fitted surrogates (descent.jl:187-260):
    # Jacobian at the iterate -> LP for the direction (mrbf_descent_direction_dev) ; SURVEY 8(f) rank 1
    if args.descent:
        Xc = dev.x.reshape(B, 1, N_VARS).contiguous()
        Jc = torch.empty((B, 1, K_OUT, N_VARS), dtype=torch.float64, device="cuda")
        out = None
        with torch.cuda.stream(stream):
            for _ in range(args.warmup):
                eng.eval_dev(model, Xc, None, Jc); out = eng.descent_direction_dev(Jc.view(B, K_OUT, N_VARS), dev.x, dev.glb, dev.gub, True, out)
            d0, d1, d2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            stream.synchronize()
            d0.record(stream)
            for _ in range(args.steps):
                eng.eval_dev(model, Xc, None, Jc)
            d1.record(stream)
            for _ in range(args.steps):
                out = eng.descent_direction_dev(Jc.view(B, K_OUT, N_VARS), dev.x, dev.glb, dev.gub, True, out)
            d2.record(stream)
            # Armijo backtracking along the LP direction: all <= 118 step sizes of every instance in one evaluation batch
            dn = out[0].abs().amax(dim=1).clamp_min(1e-300)
            dirn = (out[0] / dn[:, None]).contiguous()
            bt = eng.backtrack_dev(model, dev.x, dirn, dn, out[1])       # warm-up (tiles the model for the tensor-path sweep once)
            d2b = torch.cuda.Event(enable_timing=True); d2b.record(stream)
            for _ in range(args.steps):
                bt = eng.backtrack_dev(model, dev.x, dirn, dn, out[1], out=bt)
            d3 = torch.cuda.Event(enable_timing=True); d3.record(stream)
            stream.synchronize()
        jac_ms, lp_ms = d0.elapsed_time(d1) / args.steps, d1.elapsed_time(d2) / args.steps
        bt_ms = d2b.elapsed_time(d3) / args.steps
        secondary.append({"metric": "descent_steps_per_s", "value": world * B / ((jac_ms + lp_ms + bt_ms) * 1e-3), "unit": "steps/s",
                          "ms_per_step": jac_ms + lp_ms + bt_ms, "backtrack_ms": bt_ms,
                          "mean_backtrack_index": float(bt[0].double().mean().item()),
                          "config": {"workload": f"{B} instances per GPU: Jacobian at the iterate + exact LP direction + Armijo backtracking over the "
                                                 "surrogate (119 trial points per instance in one batch), descent.jl:187-260"}})
        secondary.append({"metric": "descent_directions_per_s", "value": world * B / ((jac_ms + lp_ms) * 1e-3), "unit": "directions/s",
                          "ms_per_step": jac_ms + lp_ms, "jacobian_ms": jac_ms, "lp_ms": lp_ms,
                          "lp_iterations_mean": float(out[2].double().mean().item()), "lp_ok": int((out[3] == 0).sum().item()),
                          "config": {"workload": f"{B} instances per GPU: surrogate Jacobian at the iterate (k={K_OUT}, n={N_VARS}) + exact LP "
                                                 "for the constrained steepest-descent direction"}})


    # ---- secondary metric: Pascoletti-Serafini inner solves on the fitted surrogates (descent.jl:478-581):
    # one mrbf_ps_solve_dev per batch = 26 generations of 620 individuals per instance, every generation one batched evaluation
    if args.ps:
        r_dir = torch.ones((B, K_OUT), dtype=torch.float64, device="cuda")
        lb_e = torch.clamp(dev.x - DELTA, min=0.0).contiguous(); ub_e = torch.clamp(dev.x + DELTA, max=1.0).contiguous()
        mx0 = torch.empty((B, 1, K_OUT), dtype=torch.float64, device="cuda")
        with torch.cuda.stream(stream):
            eng.eval_dev(model, dev.x.reshape(B, 1, N_VARS).contiguous(), mx0, None)
            pso, used = eng.ps_solve_dev(model, dev.x, lb_e, ub_e, mx0.view(B, K_OUT), r_dir, K_OUT, -1, -1, -1, 1)      # warm-up (allocations)
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            stream.synchronize(); l0p = eng.launch_count
            p0.record(stream)
            reps_ps = max(1, min(3, args.steps))
            for i_ in range(reps_ps):
                pso, used = eng.ps_solve_dev(model, dev.x, lb_e, ub_e, mx0.view(B, K_OUT), r_dir, K_OUT, -1, -1, -1, 2 + i_, out=pso)
            p1.record(stream)
            stream.synchronize()
        ps_ms = p0.elapsed_time(p1) / reps_ps
        tps = torch.tensor([ps_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tps, op=dist.ReduceOp.MAX)
        ps_ms = float(tps.item())
        fe_ps = N_DB * (3 * N_VARS + 1 + 2 * K_OUT) + 2 * (N_VARS + 1) * K_OUT
        secondary.append({"metric": "ps_solves_per_s", "value": world * B / (ps_ms * 1e-3), "unit": "solves/s", "ms_per_step": ps_ms,
                          "surrogate_evals_per_s": world * B * used / (ps_ms * 1e-3), "evals_per_solve": int(used),
                          "gpu_launches_per_solve_batch": int((eng.launch_count - l0p) // reps_ps),
                          "mean_tau": float(pso[0].mean().item()), "found": int(pso[3].sum().item()),
                          "fp64_frac_of_peak_eval_only": world * B * used / (ps_ms * 1e-3) / world * fe_ps / 1e12 / peak,
                          "config": {"workload": f"{B} instances per GPU (the C3 models: n={N_VARS}, k={K_OUT}, {N_DB} centres): Pascoletti-Serafini problem in the "
                                                 f"trust region box, direction r = 1, population 20 (n + 1) = {20 * (N_VARS + 1)}, budget 500 (n + 1) evaluations "
                                                 "(the reference's NLopt :GN_ISRES defaults, descent.jl:373, 418)"}})

    # ---- secondary metric: a heterogeneous batch (half of every database inside the trust region, mixed radii and budgets) -- rounds 1,
    # 2 and 3 all take part, instances differ in their control flow -- and a descent step on it with a step length that makes Armijo backtrack
    if args.hetero:
        host_h = synthetic.multistart_batch(B, n=N_VARS, n_db=N_DB, delta=DELTA, delta_max=DELTA_MAX, func=synthetic.zdt3,
                                            first_instance=rank * B, local_fraction=0.5)
        rng_h = np.random.default_rng(12345 + rank)
        host_h["delta"] = DELTA * rng_h.choice([0.25, 0.5, 1.0, 2.0], size=B)
        host_h["max_new"] = rng_h.choice([0, 2, 2**31 - 1], size=B, p=[0.05, 0.05, 0.9]).astype(np.int32)   # an exhausted budget is the rare case
        host_h["flags_in"][:, 0] = rng_h.integers(0, 2, size=B)
        dev_h = upload_batch(host_h, f"cuda:{local}")
        b_h = MultistartBuilder(eng, cfg, DELTA_MAX)
        with torch.cuda.stream(stream):
            m_h = None
            for _ in range(args.warmup):
                m_h, sel_h, st_h = b_h.step(dev_h, recycle=m_h); stream.synchronize()
            h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            h0.record(stream)
            for _ in range(args.steps):
                m_h, sel_h, st_h = b_h.step(dev_h, recycle=m_h)
            h1.record(stream)
            stream.synchronize()
            h_ms = h0.elapsed_time(h1) / args.steps
            eng.profile_enable(True)
            m_h, sel_h, st_h = b_h.step(dev_h, recycle=m_h)
            prof_h = eng.profile_read()
            eng.profile_enable(False)
            # descent on it: Jacobian, LP, Armijo from a step of 4 Delta (longer than the model is good for: the loop has to shrink)
            Xh = dev_h.x.reshape(B, 1, N_VARS).contiguous()
            Jh = torch.empty((B, 1, K_OUT, N_VARS), dtype=torch.float64, device="cuda")
            eng.eval_dev(m_h, Xh, None, Jh)
            oh = eng.descent_direction_dev(Jh.view(B, K_OUT, N_VARS), dev_h.x, dev_h.glb, dev_h.gub, True, None)
            dnh = oh[0].abs().amax(dim=1).clamp_min(1e-300)
            dirh = (oh[0] / dnh[:, None]).contiguous()
            bth = eng.backtrack_dev(m_h, dev_h.x, dirh, torch.minimum(dnh, 4.0 * dev_h.delta).contiguous(), oh[1])
            stream.synchronize()
        th = torch.tensor([h_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(th, op=dist.ReduceOp.MAX)
        h_ms = float(th.item())
        cnt = {a: float(getattr(sel_h, a).double().mean().item()) for a in ("n_r1", "n_r2", "n_r3", "n_r4")}
        secondary.append({"metric": METRIC, "variant": "heterogeneous", "value": world * B / (h_ms * 1e-3), "unit": UNIT, "ms_per_step": h_ms,
                          "builds_ok": int((st_h == 0).sum().item()), "mean_round_sizes": cnt,
                          "kernel_ms": {k_: round(v_, 4) for k_, v_ in prof_h.items() if k_ != "eval"},
                          "fully_linear_fraction": float(sel_h.flags_out[:, 0].double().mean().item()),
                          "rebuilt_fraction": float(sel_h.flags_out[:, 1].double().mean().item()),
                          "descent_on_it": {"lp_iterations_mean": float(oh[2].double().mean().item()), "lp_ok": int((oh[3] == 0).sum().item()),
                                            "mean_backtrack_index": float(bth[0].double().mean().item()),
                                            "backtracked_fraction": float((bth[0] > 0).double().mean().item())},
                          "config": {"workload": f"{B} instances per GPU, database snapshots of {N_DB} sites with half of them inside the trust region, radii "
                                                 "Delta x {1/4, 1/2, 1, 2}, round-3 budgets {0, 2, unlimited} with probabilities {5, 5, 90} % (budget-limited instances start round 4 under-poised: the literal kernel walks them until they are poised, then the register kernels continue), ensure_fully_linear on for half of the instances"}})
        m_h.free(); del dev_h, b_h
        # the LP alone on a workload where the simplex has to pivot: four conflicting random gradients, iterates partly on the bounds
        g_l = torch.Generator(device="cuda"); g_l.manual_seed(7 + rank)
        Jr = torch.randn((B, 4, N_VARS), dtype=torch.float64, device="cuda", generator=g_l)
        xr = torch.rand((B, N_VARS), dtype=torch.float64, device="cuda", generator=g_l)
        xr = torch.where(torch.rand((B, N_VARS), device="cuda", generator=g_l) < 0.2, torch.round(xr), xr).contiguous()
        with torch.cuda.stream(stream):
            ol = eng.descent_direction_dev(Jr, xr, dev.glb, dev.gub, True, None)
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record(stream)
            for _ in range(args.steps):
                ol = eng.descent_direction_dev(Jr, xr, dev.glb, dev.gub, True, ol)
            q1.record(stream)
            stream.synchronize()
        lp_ms2 = q0.elapsed_time(q1) / args.steps
        secondary.append({"metric": "descent_lp_per_s", "value": world * B / (lp_ms2 * 1e-3), "unit": "LPs/s", "ms_per_step": lp_ms2,
                          "lp_iterations_mean": float(ol[2].double().mean().item()), "lp_iterations_max": int(ol[2].max().item()),
                          "lp_ok": int((ol[3] == 0).sum().item()), "mean_omega": float(ol[1].mean().item()),
                          "config": {"workload": f"{B} LPs per GPU: n={N_VARS}, k=4 random normal gradients, 20 % of the coordinates of the iterate on a bound"}})

    # ---- secondary metric: config C4 (n = 200 convex quadratics; 5 RBF objectives in one group + a nonlinear constraint g(x) = sum x^2 - r^2
    # as a second RBF group, large_scale_benchmarks.jl:154-160 settings: cubic, 2n + 1 = 401 model points, theta_1 = 2, theta_pivot = 1/4):
    # per instance and iteration two surrogate updates (SurrogateContainer.jl:334-391), then the Jacobians / values the steepest-descent
    # step and compute_normal_step read (descent.jl:196, 705-708).  One wave of instances (one per SM).
    if args.c4_instances > 0:
        n4, k4, ndb4, B4 = 200, 5, 400, args.c4_instances
        rng4 = np.random.default_rng(4 + rank)
        a4 = rng4.random((k4, n4)); D4 = 0.5 + rng4.random((k4, n4))
        x4 = 0.3 + 0.4 * rng4.random((B4, n4))
        s4 = np.zeros((B4, ndb4, n4)); s4[:, 0] = x4
        rad4 = rng4.random((B4, ndb4 - 1, 1)) ** 0.25
        s4[:, 1:] = np.clip(x4[:, None, :] + (rng4.random((B4, ndb4 - 1, n4)) * 2 - 1) * 0.4 * rad4, 0.0, 1.0)
        v_obj = np.stack([np.sum(D4[j] * (s4 - a4[j]) ** 2, axis=-1) for j in range(k4)], axis=-1)
        v_con = (np.sum(s4 ** 2, axis=-1) - 0.33 * n4)[..., None]
        common = dict(n_db=np.full(B4, ndb4, np.int32), x_index=np.ones(B4, np.int32), x=x4, delta=np.full(B4, 0.1), glb=np.zeros(n4),
                      gub=np.ones(n4), flags_in=np.tile(np.array([[1, 0]], np.int32), (B4, 1)), max_new=np.full(B4, 2**31 - 1, np.int32))
        cfg_o = mb.RbfConfig(kernel="cubic", max_model_points=2 * n4 + 1, theta_enlarge_1=2.0, theta_pivot=0.25)
        cfg_c = mb.RbfConfig(kernel="multiquadric", max_model_points=2 * n4 + 1, theta_enlarge_1=2.0, theta_pivot=0.25)
        dev_o = upload_batch(dict(sites=s4, values=v_obj, **common), f"cuda:{local}")
        dev_c = upload_batch(dict(sites=s4, values=v_con, **common), f"cuda:{local}")
        b_o, b_c = MultistartBuilder(eng, cfg_o, DELTA_MAX), MultistartBuilder(eng, cfg_c, DELTA_MAX)
        X4 = dev_o.x.reshape(B4, 1, n4).contiguous()
        Jo = torch.empty((B4, 1, k4, n4), dtype=torch.float64, device="cuda"); Jcn = torch.empty((B4, 1, 1, n4), dtype=torch.float64, device="cuda")
        Yo = torch.empty((B4, 1, k4), dtype=torch.float64, device="cuda"); Ycn = torch.empty((B4, 1, 1), dtype=torch.float64, device="cuda")
        with torch.cuda.stream(stream):
            m_o = m_c = None
            for _ in range(2):
                m_o, sel_o, st_o = b_o.step(dev_o, recycle=m_o); m_c, sel_c, st_c = b_c.step(dev_c, recycle=m_c)
            c0, c1, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            stream.synchronize()
            reps4 = max(1, min(3, args.steps))
            c0.record(stream)
            for _ in range(reps4):
                m_o, sel_o, st_o = b_o.step(dev_o, recycle=m_o); m_c, sel_c, st_c = b_c.step(dev_c, recycle=m_c)
            c1.record(stream)
            for _ in range(reps4):
                eng.eval_dev(m_o, X4, Yo, Jo); eng.eval_dev(m_c, X4, Ycn, Jcn)
                o4 = eng.descent_direction_dev(Jo.view(B4, k4, n4), dev_o.x, dev_o.glb, dev_o.gub, True, None)
            c2.record(stream)
            stream.synchronize()
        ms4, ms4d = c0.elapsed_time(c1) / reps4, c1.elapsed_time(c2) / reps4
        same_r123 = bool(torch.equal(sel_o.r1, sel_c.r1) and torch.equal(sel_o.r2, sel_c.r2) and torch.equal(sel_o.n_r3, sel_c.n_r3))
        secondary.append({"metric": "c4_instance_updates_per_s", "value": world * B4 / (ms4 * 1e-3), "unit": "instance updates/s (2 groups each)",
                          "ms_per_step": ms4, "builds_ok": [int((st_o == 0).sum().item()), int((st_c == 0).sum().item())],
                          "mean_training_points": [float((1 + sel_o.n_r1 + sel_o.n_r2 + sel_o.n_r3 + sel_o.n_r4).double().mean().item()),
                                                   float((1 + sel_c.n_r1 + sel_c.n_r2 + sel_c.n_r3 + sel_c.n_r4).double().mean().item())],
                          "rounds_1_3_equal_across_groups": same_r123, "jacobians_values_lp_ms": ms4d, "lp_ok": int((o4[3] == 0).sum().item()),
                          "config": {"workload": f"C4: {B4} instances per GPU, n=200, group 1 = 5 cubic RBF objectives, group 2 = 1 multiquadric RBF constraint "
                                                 "(equal signature: rounds 1-3 coincide, test/rbf_models.jl:158-162), 401 model points, 400-site snapshots"}})
        m_o.free(); m_c.free(); del dev_o, dev_c, b_o, b_c

    # ---- secondary metric: config C2 (one optimize run: ZDT1 n = 30, k = 2, multiquadric; 61 and 496 model points): B = 1 latencies through
    # the host-pointer C ABI (what the Julia shim calls), wall clock medians
    if args.c2 and rank == 0:
        def med_ms(f, reps=9):
            f(); ts = []
            for _ in range(reps):
                t0_ = time.perf_counter(); f(); ts.append(time.perf_counter() - t0_)
            return float(np.median(ts)) * 1e3
        rows2 = []
        for ndb2, mmp2 in ((128, 61), (600, -1)):
            cfg2 = mb.RbfConfig(kernel="multiquadric", max_model_points=mmp2)
            h2 = synthetic.multistart_batch(1, n=N_VARS, n_db=ndb2, delta=DELTA, func=synthetic.zdt1)
            kept2 = [None, None]
            def sel2():
                kept2[0], kept2[1] = eng.select_points_keep(cfg2, h2["sites"], h2["n_db"], h2["x_index"], h2["x"], h2["delta"], h2["delta_max"],
                                                            h2["glb"], h2["gub"], prepared=kept2[1])
            sel2()
            r3v2 = synthetic.zdt1(kept2[0].r3_sites[0])[None]
            m2 = [None]
            def bld2():
                m2[0], _ = eng.build_prepared(cfg2, kept2[1], h2["sites"], h2["values"], h2["x_index"], kept2[0], r3v2, recycle=m2[0])
            t_s, t_b = med_ms(sel2), med_ms(bld2)
            x2 = h2["x"][:, None, :]
            t_j = med_ms(lambda: eng.eval(m2[0], x2, True, True), 21)
            J2 = eng.eval(m2[0], x2, False, True)[1][:, 0]
            t_lp = med_ms(lambda: eng.descent_direction(J2, h2["x"], h2["glb"], h2["gub"], True), 21)
            d2 = np.ones((1, N_VARS)) / np.sqrt(N_VARS)
            t_bt = med_ms(lambda: eng.backtrack(m2[0], h2["x"], d2, 0.1, 0.5), 15)
            mx2 = eng.eval(m2[0], x2, True, False)[0][:, 0]
            lb2_, ub2_ = np.maximum(0.0, h2["x"] - DELTA), np.minimum(1.0, h2["x"] + DELTA)
            t_ps = med_ms(lambda: eng.ps_solve(m2[0], h2["x"], lb2_, ub2_, mx2, np.ones((1, K_OUT)), K_OUT, -1, -1, -1, 5), 3)
            Ntr2 = int(1 + kept2[0].n_r1[0] + kept2[0].n_r2[0] + kept2[0].n_r3[0] + kept2[0].n_r4[0])
            rows2.append({"db_sites": ndb2, "max_model_points": mmp2, "training_points": Ntr2,
                          "ms": {"prepare_update_model": t_s, "update_model": t_b, "values_and_jacobian_1pt": t_j, "descent_lp": t_lp,
                                 "backtrack_118pts": t_bt, "pascoletti_serafini_solve_16120_evals": t_ps}})
            m2[0].free(); kept2[1].free()
        secondary.append({"metric": "c2_single_instance_latency_ms", "value": rows2[0]["ms"]["prepare_update_model"] + rows2[0]["ms"]["update_model"],
                          "unit": "ms per model update (61 model points)", "rows": rows2,
                          "config": {"workload": "C2: one ZDT1 n=30 k=2 instance, multiquadric; per-call wall-clock medians of the host-pointer entry points "
                                                 "(B = 1: launch-latency-bound by construction, SURVEY 7 'hard parts')"}})

    # ---- secondary metric: the whole multistart run in lock-step with device-resident databases:
    # every instance starts from its Halton point with an empty database and runs iterate! (algorithm.jl:615-917) until it stops
    if args.lockstep_iters > 0:
        import time as _time
        from morbit_jl_b200 import lockstep as LS
        x0 = synthetic.halton(rank * B + B, N_VARS)[rank * B:]
        with torch.cuda.stream(stream):
            scratch = {}
            mk = lambda: LS.LockstepDriver(cfg, synthetic.zdt3, x0, np.zeros(N_VARS), np.ones(N_VARS), LS.AlgorithmConfig(max_iter=args.lockstep_iters),
                                           device=f"cuda:{local}", capacity=N_DB, engine=eng, scratch=scratch)
            mk().run()                      # warm-up run: allocates the per-size scratch (kept factorisations, model batches) once
            barrier()
            t0 = _time.perf_counter()
            drv = mk()
            l0 = eng.launch_count
            drv.run()
            stream.synchronize()
            t_run = _time.perf_counter() - t0
        it_total = float(drv.iters_done.sum().item()); ev_total = float(drv.num_evals.sum().item())
        t = torch.tensor([t_run], dtype=torch.float64, device="cuda"); c = torch.tensor([it_total, ev_total], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(c, op=dist.ReduceOp.SUM)
        rc, cnt = np.unique(drv.ret.cpu().numpy(), return_counts=True)
        secondary.append({"metric": "multistart_instance_iterations_per_s", "value": float(c[0].item()) / float(t.item()), "unit": "iterations/s",
                          "wall_s": float(t.item()), "lockstep_iterations": int(drv.iter_counter - 1), "instance_iterations": float(c[0].item()),
                          "true_function_evaluations": float(c[1].item()), "host_function_calls": int(drv.n_func_calls),
                          "gpu_launches": int(eng.launch_count - l0),
                          "stop_codes_rank0": {LS_CODES.get(int(a), str(int(a))): int(b_) for a, b_ in zip(rc, cnt)},
                          "build_failures_rank0": int((drv.build_failures > 0).sum().item()),
                          "config": {"workload": f"C3 end to end: {B} ZDT3 n={N_VARS} instances per GPU from their Halton starting points, empty databases, "
                                                 f"max_iter={args.lockstep_iters}, database capacity {N_DB}; wall clock of the whole run incl. initialisation, the host "
                                                 "objective function (NumPy) and its copies (one warm-up run before it allocates the scratch buffers); "
                                                 "databases never leave the device"}})
        del drv

    clocks = sampler.stop()      # sampled from the first timed step to the end of the device work (build loop, e2e loop, sweeps)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import c_oracle as CO
        from oracle.rbf_oracle import RbfConfig as OracleCfg
        nthr = host_threads()
        v, sample, reps, ok_cpu = cpu_baseline(CO, OracleCfg(kernel=KERNEL), host, nthr)
        cpu = {"value": v, "unit": UNIT, "cores": nthr, "kind": "port", "builds_ok": ok_cpu,
               "sample": f"all {sample} instances of the same batch x {reps} passes, one instance per thread, one C call per pass "
                         "(orc_select_and_build_batched: rounds 1-4 + gather + solve inside the C port, oracle/rbf_oracle.c -- cheaper than the "
                         "reference's dense O(N^3) round-4 update); not Julia"}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "instances_per_gpu": B, "n_vars": N_VARS, "n_outputs": K_OUT, "db_sites": N_DB,
                           "kernel": KERNEL, "mean_training_points": float(Ntrain.mean()), "builds_ok": ok,
                           "l2": "per-step working set (database sites 126 MB + kept factorisations 0.8 GB) > 126 MB L2; no flush needed",
                           "parallelism": f"instances sharded over {world} rank(s), no data-path collective"},
                "clocks": clocks, "gpu_launches": int(launches),
                # headline end-to-end figure: device-resident databases that grow by one evaluated site per instance and step (the
                # rows an iteration of optimize appends); e2e_snapshot re-uploads the whole database snapshot every step (what a
                # caller pays that keeps its databases on the host), e2e_resident uploads no database rows at all
                "e2e": e2e_append,
                "e2e_snapshot": {"value": world * B / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": e2e_ms, "chunks": args.e2e_chunks, "builds_ok": e2e_status_ok,
                        "buffers": args.e2e_buffers, "outputs": args.e2e_outputs,
                        "path": "pinned host database snapshot -> H2D -> mrbf_select_points_keep_dev + mrbf_build_prepared_dev -> D2H of "
                                "indices/flags/status (models stay device-resident handles, as in the ABI); multistart.HostPipeline: "
                                f"{args.e2e_chunks} slice(s) of the batch, {args.e2e_buffers} device buffer(s) per slice -- the upload of the next "
                                f"snapshot runs behind the kernels of the current one; {args.e2e_outputs} set(s) of result buffers -- a step's result copy "
                                "runs behind the next step's kernels, the timed region ends when the last copy has landed"},
                "e2e_resident": e2e_resident, "strong_scaling": strong,
                "roofline": roofline, "cpu_baseline": cpu, "secondary": secondary,
                "gathered_rows": None if allrows is None else int(allrows.shape[0]), "gather_c_abi": gather_c_abi}
        emit(line)
    model.free()
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """Reference arm: the reference's own CPU formulation of the path.  The reference is Julia and cannot run in this image, so
    this times the oracle's C port (the one place besides `cpu_baseline` where bench.py executes oracle/) with all host threads on
    the SAME workload as the GPU arm: every step is one pass over all `--instances` (4096) instances.  Nothing of the product is
    imported here (no libmorbit_rbf.so in this process)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    synthetic = load_synthetic()
    from oracle import c_oracle as CO
    from oracle.rbf_oracle import RbfConfig as OracleCfg
    nthr = host_threads()
    cfg = OracleCfg(kernel=KERNEL)
    B = args.instances
    host = synthetic.multistart_batch(B, n=N_VARS, n_db=N_DB, delta=DELTA, delta_max=DELTA_MAX, func=synthetic.zdt3)
    times, ok = [], 0
    for i in range(args.warmup + args.steps):
        t, ok = cpu_step(CO, cfg, host, nthr)
        if i >= args.warmup:
            times.append(t)
    t = float(np.mean(times))
    v = B / t
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "instances_per_gpu": B, "n_vars": N_VARS, "n_outputs": K_OUT, "db_sites": N_DB,
                       "kernel": KERNEL, "builds_ok": ok,
                       "note": "one host runs this arm whatever --gpus says: at N > 1 the GPU arm processes N x this batch"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": nthr, "kind": "port",
                             "sample": f"all {B} instances per step, one instance per thread, one C call per step (rounds 1-4 + gather + "
                                       "solve inside oracle/rbf_oracle.c); C port of the reference path (no Julia in this image)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the process's original stdout; everything else that writes to fd 1 (NCCL's version
    banner, library chatter) has been redirected to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--instances", type=int, default=B_PER_GPU, help="instances per GPU (default: the C3 workload)")
    ap.add_argument("--eval-points", type=int, default=10**6, help="C5 trial points for the secondary metric (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-descent", dest="descent", action="store_false", help="skip the steepest-descent secondary metric")
    ap.add_argument("--no-ps", dest="ps", action="store_false", help="skip the Pascoletti-Serafini secondary metric")
    ap.add_argument("--no-hetero", dest="hetero", action="store_false", help="skip the heterogeneous-batch secondary metric")
    ap.add_argument("--c4-instances", type=int, default=148, help="instances of the C4 two-group secondary metric (0 = skip)")
    ap.add_argument("--no-c2", dest="c2", action="store_false", help="skip the C2 single-instance latency secondary")
    ap.add_argument("--lockstep-iters", type=int, default=20, help="max_iter of the lock-step multistart run (secondary metric; 0 = skip)")
    ap.add_argument("--e2e-chunks", type=int, default=1, help="slices of the batch in the end-to-end pipeline")
    ap.add_argument("--e2e-outputs", type=int, default=2, help="sets of result buffers (2: a step's result copy runs behind the next step's kernels)")
    ap.add_argument("--e2e-buffers", type=int, default=2, help="device buffers per slice (2: the next step's upload overlaps this step's kernels)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
