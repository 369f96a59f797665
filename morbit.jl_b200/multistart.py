"""Lock-step batch of independent multistart instances (BASELINE config C3) and its sharding over GPUs.

The reference runs many `optimize` calls concurrently under Threads.@threads
(examples/large_scale_benchmarks.jl:253); instances share nothing, so they shard over ranks with no
collective on the data path.  Only the per-instance results are gathered at the end (torch.distributed:
NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from .engine import Engine, ModelBatch, Prepared, SelectResult, max_model_points


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition [lo, hi) of `total` independent instances; sizes differ by at most 1."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def train_stride_for(cfg, n: int, db_stride: int) -> int:
    """Upper bound of the training-set size: min(max_points, n + 1 + #database sites)."""
    return max(n + 1, min(max_model_points(cfg, n), n + db_stride))


@dataclass
class DeviceBatch:
    """Device-resident state of B instances (torch CUDA tensors)."""
    sites: object       # B x db_stride x n   float64
    values: object      # B x db_stride x k   float64
    n_db: object        # B  int32
    x_index: object     # B  int32
    x: object           # B x n
    delta: object       # B
    glb: object         # n
    gub: object         # n
    flags_in: object    # B x 2 int32
    max_new: object     # B int32


def upload_batch(host: dict, device: str = "cuda:0", pin: bool = False) -> DeviceBatch:
    """host: dict of NumPy arrays with DeviceBatch's field names."""
    import torch
    def up(a, dt):
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dt)
        if pin:
            t = t.pin_memory()
        return t.to(device, non_blocking=pin)
    f64, i32 = torch.float64, torch.int32
    return DeviceBatch(up(host["sites"], f64), up(host["values"], f64), up(host["n_db"], i32), up(host["x_index"], i32),
                       up(host["x"], f64), up(host["delta"], f64), up(host["glb"], f64), up(host["gub"], f64),
                       up(host["flags_in"], i32), up(host["max_new"], i32))


class MultistartBuilder:
    """select (rounds 1-4) -> gather -> build for a device-resident batch; buffers are reused across steps."""

    def __init__(self, engine: Engine, cfg, delta_max: float, func=None):
        """func (optional): the objective, (M, n) sites -> (M, k) values on the host.  Round 3 can create NEW sites (RbfModel.jl:269-307)
        whose values the build needs (eval_missing!, Databases.jl:258-277): with `func` the fused step evaluates them between the two
        phases -- one host round trip, only when some instance has new sites; without it their values are taken as zero, which is only
        right for batches that create none (e.g. the headline snapshots: every direction is covered by database sites)."""
        self.engine, self.cfg, self.delta_max, self.func = engine, cfg, float(delta_max), func
        self._sel: Optional[SelectResult] = None
        self._train = None
        self._r3_values = None
        self._status = None
        self._prepared: Optional[Prepared] = None      # reused across steps (no allocation on the hot path)

    def select(self, d: DeviceBatch) -> SelectResult:
        self._sel = self.engine.select_points_dev(self.cfg, d.sites, d.n_db, d.x_index, d.x, d.delta, self.delta_max,
                                                  d.glb, d.gub, d.flags_in, d.max_new, out=self._sel)
        return self._sel

    def build(self, d: DeviceBatch, sel: SelectResult, r3_values=None, recycle=None) -> Tuple[ModelBatch, object]:
        import torch
        B, db_stride, n = d.sites.shape
        k = d.values.shape[2]
        ts = train_stride_for(self.cfg, n, db_stride)
        if r3_values is None:
            if self._r3_values is None or self._r3_values.shape != (B, n, k):
                self._r3_values = torch.zeros((B, n, k), dtype=torch.float64, device=d.sites.device)
            r3_values = self._r3_values
        if self._train is not None and self._train[0].shape != (B, ts, n):
            self._train = None
        self._train = self.engine.gather_training_dev(d.sites, d.values, d.x_index, sel, r3_values, ts, out=self._train)
        if self._status is None or self._status.shape[0] != B:
            self._status = torch.zeros(B, dtype=torch.int32, device=d.sites.device)
        model, status = self.engine.build_dev(self.cfg, self._train[0], self._train[1], self._train[2], None, self._status,
                                              recycle=recycle)
        return model, status

    def step(self, d: DeviceBatch, fused: bool = True, recycle=None) -> Tuple[ModelBatch, SelectResult, object]:
        """One build per instance.  fused=True keeps the round-4 factorisation and builds from it
        (mrbf_select_points_keep_dev + mrbf_build_prepared_dev); fused=False is the reference's two independent
        phases (rounds 1-4, then a from-scratch solve of the gathered training set).  `recycle`: the previous
        iteration's ModelBatch, replaced in place like SurrogateContainer.jl:376-382 (no allocation per step)."""
        if not fused:
            sel = self.select(d)
            model, status = self.build(d, sel, recycle=recycle)
            return model, sel, status
        import torch
        self._sel, self._prepared = self.engine.select_points_keep_dev(self.cfg, d.sites, d.n_db, d.x_index, d.x, d.delta,
                                                                       self.delta_max, d.glb, d.gub, d.flags_in, d.max_new,
                                                                       out=self._sel, prepared=self._prepared)
        prepared = self._prepared
        B, _, n = d.sites.shape
        k = d.values.shape[2]
        if self._r3_values is None or self._r3_values.shape != (B, n, k):
            self._r3_values = torch.zeros((B, n, k), dtype=torch.float64, device=d.sites.device)
        if self._status is None or self._status.shape[0] != B:
            self._status = torch.zeros(B, dtype=torch.int32, device=d.sites.device)
        if self.func is not None:
            new = torch.arange(n, device=d.sites.device)[None, :] < self._sel.n_r3[:, None]       # rows of r3_sites that are new sites
            if bool(new.any()):
                vals = np.asarray(self.func(self._sel.r3_sites[new].cpu().numpy()), dtype=np.float64).reshape(-1, k)
                self._r3_values.zero_()
                self._r3_values[new] = torch.from_numpy(vals).to(d.sites.device)
        model, status = self.engine.build_prepared_dev(self.cfg, prepared, d.sites, d.values, d.x_index, self._sel,
                                                       self._r3_values, self._status, recycle=recycle)
        return model, self._sel, status


class HostPipeline:
    """One build per instance from HOST database snapshots (what a host-language caller such as the Julia shim hands over),
    with the copies hidden behind the kernels: the batch is cut into `chunks` slices; the pinned host -> device copy of slice
    c + 1 runs on a copy stream while slice c is processed on the engine's stream, and the device -> host copy of slice c's
    indices / flags / status runs while slice c + 1 computes.  Nothing about the results changes -- instances are independent."""

    NAMES = ("sites", "values", "n_db", "x_index", "x", "delta", "flags_in", "max_new")
    OUTS = ("r1", "n_r1", "r2", "n_r2", "n_r3", "r4", "n_r4", "flags_out")

    def __init__(self, engine: Engine, cfg, delta_max: float, host: dict, device: str, compute_stream, chunks: int = 2, buffers: int = 1,
                 outputs: int = 1, resident_db: bool = False, append_rows: int = 0):
        """resident_db: the databases (sites, values) are uploaded ONCE and stay on the device (grown with
        mrbf_db_append_dev by a driver); a step then uploads only what changes between two model updates on the same database --
        iterate, radius, flags, budget -- as in the reference's criticality loop (algorithm.jl:523-612), which rebuilds the models
        for a shrinking radius on an unchanged database.
        append_rows (with resident_db): every step additionally uploads the LAST `append_rows` rows of every database (site + values: the
        evaluations an iteration adds, Databases.jl:390-401) and appends them on the device with mrbf_db_append_dev -- the resident
        databases hold the rows before them, `n_db - append_rows` travels as the per-step size; the result of a step is unchanged."""
        import torch
        self.torch = torch
        self.resident_db = bool(resident_db)
        self.engine, self.cfg, self.chunks, self.compute = engine, cfg, chunks, compute_stream
        B = host["sites"].shape[0]
        self.bounds = [shard_range(B, c, chunks) for c in range(chunks)]
        f64, i32 = torch.float64, torch.int32
        dt = dict(sites=f64, values=f64, n_db=i32, x_index=i32, x=f64, delta=f64, flags_in=i32, max_new=i32)
        self.pinned = [{k: torch.from_numpy(np.ascontiguousarray(host[k][lo:hi])).to(dt[k]).pin_memory() for k in self.NAMES}
                       for lo, hi in self.bounds]
        glb = torch.from_numpy(np.ascontiguousarray(host["glb"])).to(f64).to(device)
        gub = torch.from_numpy(np.ascontiguousarray(host["gub"])).to(f64).to(device)
        # `buffers` device copies of every slice, used in turn: with two, the host -> device copy of the NEXT step's snapshot runs
        # behind the kernels of the current step even when the batch is not cut into slices (chunks = 1: full-size launches)
        self.buffers, self._turn = max(1, int(buffers)), 0
        self.dev = [[DeviceBatch(*(torch.empty_like(pc[k], device=device) for k in ("sites", "values", "n_db", "x_index", "x", "delta")),
                                 glb, gub, torch.empty_like(pc["flags_in"], device=device), torch.empty_like(pc["max_new"], device=device))
                     for _ in range(self.buffers)] for pc in self.pinned]
        # `outputs` sets of result buffers (select outputs, kept factorisations, models, pinned host copies), used in turn: with two,
        # the device -> host copy of a step's results runs behind the kernels of the NEXT step (drain() waits for the last one)
        self.outputs, self._uturn = max(1, int(outputs)), 0
        self.builders = [[MultistartBuilder(engine, cfg, delta_max) for _ in range(self.outputs)] for _ in range(chunks)]
        self.models = [[None] * self.outputs for _ in range(chunks)]
        self.out_pinned = [[None] * self.outputs for _ in range(chunks)]
        self.copy_in, self.copy_out = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
        self.ev_in = [torch.cuda.Event() for _ in range(chunks)]
        self.ev_comp = [[torch.cuda.Event() for _ in range(self.buffers)] for _ in range(chunks)]
        self.ev_out = [[torch.cuda.Event() for _ in range(self.outputs)] for _ in range(chunks)]
        self.step_names = tuple(k for k in self.NAMES if not (self.resident_db and k in ("sites", "values")))
        self.append_rows = int(append_rows) if self.resident_db else 0
        self.new_pinned, self.new_dev, self.n_add = [], [], []
        if self.append_rows > 0:
            a = self.append_rows
            for c, (lo, hi) in enumerate(self.bounds):
                nd = np.asarray(host["n_db"][lo:hi]).astype(np.int64)
                assert np.all(nd >= a), "append_rows larger than a database"
                rows = (nd[:, None] - a + np.arange(a)[None, :])                                 # the last a rows of every database
                ns = np.take_along_axis(np.asarray(host["sites"][lo:hi]), rows[:, :, None], axis=1)
                nv = np.take_along_axis(np.asarray(host["values"][lo:hi]), rows[:, :, None], axis=1)
                self.new_pinned.append((torch.from_numpy(np.ascontiguousarray(ns)).to(f64).pin_memory(),
                                        torch.from_numpy(np.ascontiguousarray(nv)).to(f64).pin_memory()))
                self.new_dev.append([(torch.empty_like(self.new_pinned[c][0], device=device), torch.empty_like(self.new_pinned[c][1], device=device))
                                     for _ in range(self.buffers)])
                self.n_add.append(torch.full((hi - lo,), a, dtype=i32, device=device))
                self.pinned[c]["n_db"] = (self.pinned[c]["n_db"] - a).pin_memory()               # size before this step's evaluations arrive
        if self.resident_db:
            for c, pc in enumerate(self.pinned):
                for d in self.dev[c]:
                    d.sites.copy_(pc["sites"]); d.values.copy_(pc["values"])
            torch.cuda.synchronize()
        self.h2d_bytes = sum(pc[k].numel() * pc[k].element_size() for pc in self.pinned for k in self.step_names)
        self.h2d_bytes += sum(t_.numel() * t_.element_size() for pair in self.new_pinned for t_ in pair)
        self.d2h_bytes = 0

    def step(self):
        """Enqueue one pass over the whole batch; returns immediately (synchronise the compute stream to wait for it)."""
        torch = self.torch
        t = self._turn
        self._turn = (t + 1) % self.buffers
        u = self._uturn
        self._uturn = (u + 1) % self.outputs
        for c in range(self.chunks):
            with torch.cuda.stream(self.copy_in):
                self.copy_in.wait_event(self.ev_comp[c][t])       # this device buffer of the slice is free again
                for k in self.step_names:
                    getattr(self.dev[c][t], k).copy_(self.pinned[c][k], non_blocking=True)
                if self.append_rows > 0:
                    self.new_dev[c][t][0].copy_(self.new_pinned[c][0], non_blocking=True)
                    self.new_dev[c][t][1].copy_(self.new_pinned[c][1], non_blocking=True)
                self.ev_in[c].record(self.copy_in)
            with torch.cuda.stream(self.compute):
                self.compute.wait_event(self.ev_in[c])
                self.compute.wait_event(self.ev_out[c][u])        # the last result copy out of this set of output buffers has left
                if self.append_rows > 0:                          # new_result! of this step's evaluations, on the device
                    d_ = self.dev[c][t]
                    self.engine.db_append_dev(d_.sites, d_.values, d_.n_db, self.new_dev[c][t][0], self.new_dev[c][t][1], self.n_add[c])
                self.models[c][u], sel, status = self.builders[c][u].step(self.dev[c][t], recycle=self.models[c][u])
                self.ev_comp[c][t].record(self.compute)
            outs = [getattr(sel, k) for k in self.OUTS] + [status]
            if self.out_pinned[c][u] is None:
                self.out_pinned[c][u] = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in outs]
                if u == 0:
                    self.d2h_bytes += sum(o.numel() * o.element_size() for o in outs)
            with torch.cuda.stream(self.copy_out):
                self.copy_out.wait_event(self.ev_comp[c][t])
                for p, o in zip(self.out_pinned[c][u], outs):
                    p.copy_(o, non_blocking=True)
                self.ev_out[c][u].record(self.copy_out)
        if self.outputs == 1:
            self.compute.wait_event(self.ev_out[self.chunks - 1][0])     # the step ends when its last result copy has landed
        return [m[u] for m in self.models], [o[u] for o in self.out_pinned]

    def drain(self):
        """Make the compute stream wait for every result copy enqueued so far (the end of a pipelined run with outputs > 1)."""
        for c in range(self.chunks):
            for u in range(self.outputs):
                self.compute.wait_event(self.ev_out[c][u])


def gather_results(local: np.ndarray, total: int, rank: int, world: int) -> Optional[np.ndarray]:
    """Final gather of per-instance result rows to every rank (no collective on the hot path; this is the only one).

    `local` holds this rank's shard (rows shard_range(total, rank, world)).  Uses the initialised
    torch.distributed process group (NCCL for CUDA tensors, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    if world == 1 or not dist.is_initialized():
        return local
    backend = dist.get_backend()
    counts = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    width = int(np.prod(local.shape[1:])) if local.ndim > 1 else 1
    mx = max(counts)
    pad = np.zeros((mx, width), dtype=np.float64)
    pad[: counts[rank]] = local.reshape(counts[rank], width)
    t = torch.from_numpy(pad)
    if backend == "nccl":
        t = t.cuda()
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    rows = [o.cpu().numpy()[: counts[r]] for r, o in enumerate(outs)]
    return np.concatenate(rows, axis=0).reshape((total,) + tuple(local.shape[1:]))


def gather_results_c_abi(comm, local: np.ndarray, total: int) -> np.ndarray:
    """The same final gather through the C ABI (mrbf_gather: one ncclAllGather on the library's own communicator) -- what a
    Julia host calls.  `comm`: engine.Comm; `local`: this rank's rows (shard_range(total, rank, world))."""
    world, rank = comm.world, comm.rank
    counts = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    assert local.shape[0] == counts[rank]
    parts = comm.gather(local.reshape(local.shape[0], -1), max(counts))
    for r in range(world):
        assert parts[r].shape[0] == counts[r], (r, parts[r].shape, counts[r])
    return np.concatenate(parts, axis=0).reshape((total,) + tuple(local.shape[1:]))
