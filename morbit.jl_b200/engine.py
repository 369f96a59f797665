"""Thin object layer over the C ABI: one `Engine` per (thread, device), batched calls.

Host entry points take NumPy arrays (what Julia's `ccall` would pass); `*_dev` entry points take
torch CUDA tensors (device-resident data, no copies) and enqueue on the engine's stream.
torch is plumbing only (device memory + streams); all arithmetic happens in libmorbit_rbf.so.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from ._lib import MrbfCfg, MrbfError, KERNEL_IDS


def max_model_points(cfg, n: int) -> int:
    """RbfModel.jl:356."""
    return ((n + 1) * (n + 2)) // 2 if cfg.max_model_points <= 0 else cfg.max_model_points


def to_c_cfg(cfg, shape: Optional[float] = None) -> MrbfCfg:
    sp = cfg.shape_parameter if shape is None else shape
    if isinstance(sp, str):
        raise TypeError("string shape parameters must be evaluated by the caller (RbfModel.jl:135-143)")
    return MrbfCfg(KERNEL_IDS[cfg.kernel], int(cfg.polynomial_degree), float(sp), float(cfg.theta_enlarge_1),
                   float(cfg.theta_enlarge_2), float(cfg.theta_pivot), float(cfg.theta_pivot_cholesky),
                   int(cfg.max_model_points), int(bool(cfg.use_max_points)), int(bool(cfg.optimized_sampling)), 0)


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()          # torch tensor


@dataclass
class SelectResult:
    """Outputs of prepare_update_model's rounds 1-4 for a batch (ids are 1-based)."""
    r1: object; n_r1: object
    r2: object; n_r2: object
    r3_sites: object; n_r3: object
    r4: object; n_r4: object
    dirs: object; n_dirs: object
    flags_out: object
    status: object


class ModelBatch:
    """Device-resident batch of fitted RBF models (opaque mrbf_model handle)."""

    def __init__(self, engine: "Engine", handle: int):
        self.engine = engine
        self.handle = handle
        dims = (C.c_int32 * 6)()
        engine._check(engine.lib.mrbf_model_dims(handle, dims))
        self.B, self.n, self.k, self.train_stride, self.p, self.degree = (int(v) for v in dims)

    def coeffs(self):
        w = np.zeros((self.B, self.train_stride, self.k))
        lam = np.zeros((self.B, self.p, self.k))
        self.engine._check(self.engine.lib.mrbf_model_coeffs(self.engine.ctx, self.handle, _ptr(w), _ptr(lam) if self.p else None))
        return w, lam

    def free(self):
        if self.handle:
            self.engine.lib.mrbf_free_model(self.engine.ctx, self.handle)
            self.handle = None

    def __del__(self):
        try:
            if self.handle and self.engine.ctx:
                self.free()
        except Exception:
            pass


class Prepared:
    """Round-4 factorisations kept on the device (opaque mrbf_prepared handle)."""

    def __init__(self, engine: "Engine", handle: int):
        self.engine, self.handle = engine, handle

    def free(self):
        if self.handle:
            self.engine.lib.mrbf_free_prepared(self.engine.ctx, self.handle)
            self.handle = None

    def __del__(self):
        try:
            if self.handle and self.engine.ctx:
                self.free()
        except Exception:
            pass


class Comm:
    """NCCL communicator for the final gather of per-instance results (mrbf_comm_*, mrbf_gather: the only collective of the
    path).  `unique_id()` on rank 0, hand the 128 bytes to every rank (any host channel), then `Comm(device, id, rank, world)`."""

    def __init__(self, device: int, unique_id: bytes, rank: int, world: int):
        self.lib = _lib.load()
        self.rank, self.world, self.device = int(rank), int(world), int(device)
        buf = C.create_string_buffer(bytes(unique_id), 128)
        h = C.c_void_p()
        rc = self.lib.mrbf_comm_init(self.device, buf, self.rank, self.world, C.byref(h))
        if rc != 0:
            raise MrbfError(rc, "mrbf_comm_init failed: " + self.lib.mrbf_comm_last_error(None).decode(errors="replace"))
        self.handle = h

    @staticmethod
    def unique_id() -> bytes:
        lib = _lib.load()
        buf = C.create_string_buffer(128)
        rc = lib.mrbf_comm_unique_id(buf)
        if rc != 0:
            raise MrbfError(rc, "mrbf_comm_unique_id failed: " + lib.mrbf_comm_last_error(None).decode(errors="replace"))
        return buf.raw

    def gather(self, rows: np.ndarray, max_count: int):
        """rows: (count, width) float64 of this rank -> list of per-rank arrays (counts[r], width)."""
        rows = _np(rows, np.float64)
        if rows.ndim != 2:
            rows = rows.reshape(rows.shape[0], int(np.prod(rows.shape[1:], dtype=np.int64)))   # (0, w, ...) keeps its width
        count, width = rows.shape
        out = np.zeros((self.world, int(max_count), width))
        counts = np.zeros(self.world, np.int32)
        rc = self.lib.mrbf_gather(self.handle, _ptr(rows), count, width, int(max_count), _ptr(out), _ptr(counts))
        if rc != 0:
            raise MrbfError(rc, self.lib.mrbf_comm_last_error(self.handle).decode(errors="replace"))
        return [out[r, :counts[r]] for r in range(self.world)]

    def close(self):
        if self.handle:
            self.lib.mrbf_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _StreamOrderedLib:
    """The library as an Engine sees it.  `*_dev` entry points take torch tensors, which torch produces (and will consume) on ITS current
    stream; when that is not the stream the context enqueues on -- an Engine created without `stream=` runs on the library's private
    non-blocking stream -- every such call is bracketed by two event waits: the engine's stream waits for what torch has enqueued so
    far (e.g. the zero fill of freshly allocated outputs), and torch's stream waits for the call's kernels afterwards.  Nothing is
    added when both are the same stream (bench.py, lockstep).  Every other symbol is passed through untouched."""

    def __init__(self, lib, engine):
        import weakref
        self._lib, self._engine = lib, weakref.ref(engine)

    def __getattr__(self, name):
        f = getattr(self._lib, name)
        if not name.endswith("_dev"):
            return f

        def ordered(*args):
            eng = self._engine()
            ext = eng._foreign_stream() if eng is not None else None
            if ext is None:
                return f(*args)
            import torch
            cur = torch.cuda.current_stream(eng.device)
            ext.wait_stream(cur)
            rc = f(*args)
            cur.wait_stream(ext)
            return rc

        self.__dict__[name] = ordered
        return ordered


class Engine:
    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self.lib = _StreamOrderedLib(_lib.load(), self)
        ctx = C.c_void_p()
        rc = self.lib.mrbf_init(int(device), C.byref(ctx))
        if rc != 0:
            raise MrbfError(rc, f"mrbf_init(device={device}) failed: no usable CUDA device (there is no CPU fallback)")
        self.ctx = ctx
        self.device = device
        self._ext = None
        if stream is not None:
            self._check(self.lib.mrbf_set_stream(self.ctx, C.c_void_p(stream)))
        h = C.c_void_p()
        self._check(self.lib.mrbf_get_stream(self.ctx, C.byref(h)))
        self._stream = int(h.value or 0)

    def _foreign_stream(self):
        """torch's view of the context's stream when it differs from torch's current stream on this device, else None."""
        import torch
        if torch.cuda.current_stream(self.device).cuda_stream == self._stream:
            return None
        if self._ext is None:
            self._ext = torch.cuda.ExternalStream(self._stream, device=self.device)
        return self._ext

    def close(self):
        if self.ctx:
            self.lib.mrbf_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise MrbfError(rc, self.lib.mrbf_last_error(self.ctx).decode(errors="replace"))

    def sync(self):
        self._check(self.lib.mrbf_sync(self.ctx))

    def set_isapprox_rtol(self, rtol: float):
        """rtol of the reference's `Δ ≈ Δ_max` test (RbfModel.jl:588): sqrt(eps(Float64)) by default; a run with the reference's
        default algorithm config (Float32 literals) uses sqrt(eps(Float32)) -- see mrbf_set_isapprox_rtol in include/morbit_rbf.h."""
        self._check(self.lib.mrbf_set_isapprox_rtol(self.ctx, float(rtol)))

    def profile_enable(self, on: bool = True):
        self._check(self.lib.mrbf_profile_enable(self.ctx, int(on)))

    def profile_read(self) -> dict:
        ms = (C.c_double * 8)()
        self._check(self.lib.mrbf_profile_read(self.ctx, ms))
        return dict(rounds123=ms[0], round4=ms[1], gather=ms[2], build=ms[3], eval=ms[4], round4_fallback=ms[5],
                    build_prepared=ms[6], round4_prefix=ms[7])

    @property
    def launch_count(self) -> int:
        return int(self.lib.mrbf_launch_count(self.ctx))

    # ------------------------------------------------------------------ rounds 1-4 (host buffers)
    def select_points(self, cfg, sites, n_db, x_index, x, delta, delta_max, glb, gub, ensure_fully_linear=False,
                      force_rebuild=False, max_new=2**31 - 1) -> SelectResult:
        sites = _np(sites, np.float64)
        B, db_stride, n = sites.shape
        n_db = _np(np.broadcast_to(n_db, (B,)), np.int32); x_index = _np(np.broadcast_to(x_index, (B,)), np.int32)
        x = _np(x, np.float64).reshape(B, n); delta = _np(np.broadcast_to(delta, (B,)), np.float64)
        glb = _np(np.broadcast_to(glb, (n,)), np.float64); gub = _np(np.broadcast_to(gub, (n,)), np.float64)
        flags_in = _np(np.stack([np.broadcast_to(ensure_fully_linear, (B,)), np.broadcast_to(force_rebuild, (B,))], 1), np.int32)
        max_new = _np(np.clip(np.broadcast_to(max_new, (B,)), 0, 2**31 - 1), np.int32)
        mp = max_model_points(cfg, n)
        r1 = np.zeros((B, n), np.int32); r2 = np.zeros((B, n), np.int32); r4 = np.zeros((B, mp), np.int32)
        cnt = [np.zeros(B, np.int32) for _ in range(5)]
        r3 = np.zeros((B, n, n)); dirs = np.zeros((B, n, n))
        flags_out = np.zeros((B, 2), np.int32); status = np.zeros(B, np.int32)
        ccfg = to_c_cfg(cfg)
        self._check(self.lib.mrbf_select_points(
            self.ctx, C.byref(ccfg), B, n, db_stride, _ptr(sites), _ptr(n_db), _ptr(x_index), _ptr(x), _ptr(delta),
            float(delta_max), _ptr(glb), _ptr(gub), _ptr(flags_in), _ptr(max_new), _ptr(r1), _ptr(cnt[0]), _ptr(r2),
            _ptr(cnt[1]), _ptr(r3), _ptr(cnt[2]), mp, _ptr(r4), _ptr(cnt[3]), _ptr(dirs), _ptr(cnt[4]), _ptr(flags_out),
            _ptr(status)))
        return SelectResult(r1, cnt[0], r2, cnt[1], r3, cnt[2], r4, cnt[3], dirs, cnt[4], flags_out, status)

    def select_points_keep(self, cfg, sites, n_db, x_index, x, delta, delta_max, glb, gub, ensure_fully_linear=False,
                           force_rebuild=False, max_new=2**31 - 1, prepared: Optional["Prepared"] = None):
        """select_points that also keeps the round-4 factorisation on the device (host buffers in and out); returns
        (SelectResult, Prepared).  build_prepared turns the pair into the model without a from-scratch solve."""
        sites = _np(sites, np.float64)
        B, db_stride, n = sites.shape
        n_db = _np(np.broadcast_to(n_db, (B,)), np.int32); x_index = _np(np.broadcast_to(x_index, (B,)), np.int32)
        x = _np(x, np.float64).reshape(B, n); delta = _np(np.broadcast_to(delta, (B,)), np.float64)
        glb = _np(np.broadcast_to(glb, (n,)), np.float64); gub = _np(np.broadcast_to(gub, (n,)), np.float64)
        flags_in = _np(np.stack([np.broadcast_to(ensure_fully_linear, (B,)), np.broadcast_to(force_rebuild, (B,))], 1), np.int32)
        max_new = _np(np.clip(np.broadcast_to(max_new, (B,)), 0, 2**31 - 1), np.int32)
        mp = max_model_points(cfg, n)
        r1 = np.zeros((B, n), np.int32); r2 = np.zeros((B, n), np.int32); r4 = np.zeros((B, mp), np.int32)
        cnt = [np.zeros(B, np.int32) for _ in range(5)]
        r3 = np.zeros((B, n, n)); dirs = np.zeros((B, n, n))
        flags_out = np.zeros((B, 2), np.int32); status = np.zeros(B, np.int32)
        ccfg = to_c_cfg(cfg)
        handle = C.c_void_p(prepared.handle if prepared is not None else None)
        if prepared is not None:
            prepared.handle = None
        self._check(self.lib.mrbf_select_points_keep(
            self.ctx, C.byref(ccfg), B, n, db_stride, _ptr(sites), _ptr(n_db), _ptr(x_index), _ptr(x), _ptr(delta),
            float(delta_max), _ptr(glb), _ptr(gub), _ptr(flags_in), _ptr(max_new), _ptr(r1), _ptr(cnt[0]), _ptr(r2),
            _ptr(cnt[1]), _ptr(r3), _ptr(cnt[2]), mp, _ptr(r4), _ptr(cnt[3]), _ptr(dirs), _ptr(cnt[4]), _ptr(flags_out),
            _ptr(status), C.byref(handle)))
        res = SelectResult(r1, cnt[0], r2, cnt[1], r3, cnt[2], r4, cnt[3], dirs, cnt[4], flags_out, status)
        if prepared is not None:
            prepared.handle = handle.value
            return res, prepared
        return res, Prepared(self, handle.value)

    def build_prepared(self, cfg, prepared: "Prepared", sites, values, x_index, sel: SelectResult, r3_values=None,
                       recycle: Optional[ModelBatch] = None, raise_on_failure: bool = True):
        """update_model from the factorisation kept by select_points_keep (host buffers).  sites / values: the same database
        arrays (B x db_stride x n / k) the selection saw; r3_values: B x n x k values of the new round-3 sites."""
        sites = _np(sites, np.float64); values = _np(values, np.float64)
        B, db_stride, n = sites.shape
        k = values.shape[2]
        x_index = _np(np.broadcast_to(x_index, (B,)), np.int32)
        r3v = None if r3_values is None else _np(r3_values, np.float64).reshape(B, n, k)
        status = np.zeros(B, np.int32)
        handle = C.c_void_p(recycle.handle if recycle is not None else None)
        if recycle is not None:
            recycle.handle = None
        ccfg = to_c_cfg(cfg)
        rc = self.lib.mrbf_build_prepared(
            self.ctx, C.byref(ccfg), prepared.handle, k, _ptr(sites), _ptr(values), _ptr(_np(sel.r3_sites, np.float64)), _ptr(r3v),
            _ptr(x_index), _ptr(_np(sel.r1, np.int32)), _ptr(_np(sel.n_r1, np.int32)), _ptr(_np(sel.r2, np.int32)),
            _ptr(_np(sel.n_r2, np.int32)), _ptr(_np(sel.n_r3, np.int32)), C.byref(handle), _ptr(status))
        # MRBF_ENUMERIC returns a VALID handle (some instances failed numerically, status[] says which); every other error
        # has already released it inside the library (include/morbit_rbf.h, ownership rule of mrbf_build*)
        if rc != 0 and not (rc == _lib.MRBF_ENUMERIC and not raise_on_failure):
            if rc == _lib.MRBF_ENUMERIC and handle.value:
                self.lib.mrbf_free_model(self.ctx, handle)
            self._check(rc)
        return ModelBatch(self, handle.value), status

    def round4(self, cfg, sites, n_db, lb2, ub2, found, n_found, extra_sites=None, n_extra=None):
        sites = _np(sites, np.float64)
        B, db_stride, n = sites.shape
        n_db = _np(np.broadcast_to(n_db, (B,)), np.int32)
        lb2 = _np(np.broadcast_to(lb2, (B, n)), np.float64); ub2 = _np(np.broadcast_to(ub2, (B, n)), np.float64)
        found = _np(found, np.int32).reshape(B, -1); n_found = _np(np.broadcast_to(n_found, (B,)), np.int32)
        es = 0
        if extra_sites is not None:
            extra_sites = _np(extra_sites, np.float64).reshape(B, -1, n); es = extra_sites.shape[1]
            n_extra = _np(np.broadcast_to(n_extra, (B,)), np.int32)
        mp = max_model_points(cfg, n)
        r4 = np.zeros((B, max(mp, 1)), np.int32); n_r4 = np.zeros(B, np.int32); status = np.zeros(B, np.int32)
        ccfg = to_c_cfg(cfg)
        self._check(self.lib.mrbf_round4(self.ctx, C.byref(ccfg), B, n, db_stride, _ptr(sites), _ptr(n_db), _ptr(lb2), _ptr(ub2),
                                         found.shape[1], _ptr(found), _ptr(n_found), es,
                                         _ptr(extra_sites) if es else None, _ptr(n_extra) if es else None,
                                         r4.shape[1], _ptr(r4), _ptr(n_r4), _ptr(status)))
        return r4, n_r4, status

    # ------------------------------------------------------------------ rounds 1-4 (device tensors)
    def select_points_dev(self, cfg, sites, n_db, x_index, x, delta, delta_max, glb, gub, flags_in, max_new, out=None):
        """All arguments are torch CUDA tensors (float64 / int32); returns a SelectResult of CUDA tensors."""
        import torch
        B, db_stride, n = sites.shape
        mp = max(1, min(max_model_points(cfg, n), db_stride))    # round 4 accepts database sites only: never more than db_stride ids
        dev = sites.device
        if out is None:
            i32 = dict(dtype=torch.int32, device=dev); f64 = dict(dtype=torch.float64, device=dev)
            out = SelectResult(torch.zeros((B, n), **i32), torch.zeros(B, **i32), torch.zeros((B, n), **i32), torch.zeros(B, **i32),
                               torch.zeros((B, n, n), **f64), torch.zeros(B, **i32), torch.zeros((B, mp), **i32),
                               torch.zeros(B, **i32), torch.zeros((B, n, n), **f64), torch.zeros(B, **i32),
                               torch.zeros((B, 2), **i32), torch.zeros(B, **i32))
        ccfg = to_c_cfg(cfg)
        self._check(self.lib.mrbf_select_points_dev(
            self.ctx, C.byref(ccfg), B, n, db_stride, _ptr(sites), _ptr(n_db), _ptr(x_index), _ptr(x), _ptr(delta),
            float(delta_max), _ptr(glb), _ptr(gub), _ptr(flags_in), _ptr(max_new), _ptr(out.r1), _ptr(out.n_r1), _ptr(out.r2),
            _ptr(out.n_r2), _ptr(out.r3_sites), _ptr(out.n_r3), out.r4.shape[1], _ptr(out.r4), _ptr(out.n_r4), _ptr(out.dirs),
            _ptr(out.n_dirs), _ptr(out.flags_out), _ptr(out.status)))
        return out

    def select_points_keep_dev(self, cfg, sites, n_db, x_index, x, delta, delta_max, glb, gub, flags_in, max_new, out=None,
                               prepared=None):
        """select_points_dev that also keeps the round-4 factorisation on the device; returns (SelectResult, Prepared)."""
        import torch
        B, db_stride, n = sites.shape
        mp = max(1, min(max_model_points(cfg, n), db_stride))    # round 4 accepts database sites only: never more than db_stride ids
        dev = sites.device
        if out is None:
            i32 = dict(dtype=torch.int32, device=dev); f64 = dict(dtype=torch.float64, device=dev)
            out = SelectResult(torch.zeros((B, n), **i32), torch.zeros(B, **i32), torch.zeros((B, n), **i32), torch.zeros(B, **i32),
                               torch.zeros((B, n, n), **f64), torch.zeros(B, **i32), torch.zeros((B, mp), **i32),
                               torch.zeros(B, **i32), torch.zeros((B, n, n), **f64), torch.zeros(B, **i32),
                               torch.zeros((B, 2), **i32), torch.zeros(B, **i32))
        ccfg = to_c_cfg(cfg)
        handle = C.c_void_p(prepared.handle if prepared is not None else None)
        if prepared is not None:
            prepared.handle = None          # ownership moves through the call (the library may replace the handle)
        self._check(self.lib.mrbf_select_points_keep_dev(
            self.ctx, C.byref(ccfg), B, n, db_stride, _ptr(sites), _ptr(n_db), _ptr(x_index), _ptr(x), _ptr(delta),
            float(delta_max), _ptr(glb), _ptr(gub), _ptr(flags_in), _ptr(max_new), _ptr(out.r1), _ptr(out.n_r1), _ptr(out.r2),
            _ptr(out.n_r2), _ptr(out.r3_sites), _ptr(out.n_r3), out.r4.shape[1], _ptr(out.r4), _ptr(out.n_r4), _ptr(out.dirs),
            _ptr(out.n_dirs), _ptr(out.flags_out), _ptr(out.status), C.byref(handle)))
        if prepared is not None:
            prepared.handle = handle.value
            return out, prepared
        return out, Prepared(self, handle.value)

    def build_prepared_dev(self, cfg, prepared, sites, values, x_index, sel: SelectResult, r3_values=None, status=None,
                           recycle: Optional[ModelBatch] = None):
        """update_model from the factorisation kept by select_points_keep_dev (general route for the rest).
        `recycle`: an earlier ModelBatch whose device buffers are reused when the shapes match (it must not be used
        afterwards; the returned ModelBatch replaces it)."""
        import torch
        B = sites.shape[0]
        k = values.shape[2]
        if status is None:
            status = torch.zeros(B, dtype=torch.int32, device=sites.device)
        handle = C.c_void_p(recycle.handle if recycle is not None else None)
        if recycle is not None:
            recycle.handle = None               # ownership moves through the call
        ccfg = to_c_cfg(cfg)
        self._check(self.lib.mrbf_build_prepared_dev(
            self.ctx, C.byref(ccfg), prepared.handle, k, _ptr(sites), _ptr(values), _ptr(sel.r3_sites), _ptr(r3_values),
            _ptr(x_index), _ptr(sel.r1), _ptr(sel.n_r1), _ptr(sel.r2), _ptr(sel.n_r2), _ptr(sel.n_r3), C.byref(handle), _ptr(status)))
        return ModelBatch(self, handle.value), status

    def gather_training_dev(self, sites, values, x_index, sel: SelectResult, r3_values, train_stride, out=None):
        import torch
        B, db_stride, n = sites.shape
        k = values.shape[2]
        dev = sites.device
        if out is None:
            out = (torch.zeros((B, train_stride, n), dtype=torch.float64, device=dev),
                   torch.zeros((B, train_stride, k), dtype=torch.float64, device=dev),
                   torch.zeros(B, dtype=torch.int32, device=dev))
        self._check(self.lib.mrbf_gather_training_dev(
            self.ctx, B, n, k, db_stride, _ptr(sites), _ptr(values), _ptr(x_index), _ptr(sel.r1), _ptr(sel.n_r1), _ptr(sel.r2),
            _ptr(sel.n_r2), _ptr(sel.r3_sites), _ptr(r3_values), _ptr(sel.n_r3), sel.r4.shape[1], _ptr(sel.r4), _ptr(sel.n_r4),
            train_stride, _ptr(out[0]), _ptr(out[1]), _ptr(out[2])))
        return out

    # ------------------------------------------------------------------ build
    def build(self, cfg, sites, values, N, shape=None, raise_on_failure: bool = True):
        sites = _np(sites, np.float64); values = _np(values, np.float64)
        B, ts, n = sites.shape
        k = values.shape[2]
        N = _np(np.broadcast_to(N, (B,)), np.int32)
        shape_a = None if shape is None else _np(np.broadcast_to(shape, (B,)), np.float64)
        status = np.zeros(B, np.int32)
        handle = C.c_void_p()
        ccfg = to_c_cfg(cfg)
        rc = self.lib.mrbf_build(self.ctx, C.byref(ccfg), B, n, k, ts, _ptr(N), _ptr(sites), _ptr(values), _ptr(shape_a),
                                 C.byref(handle), _ptr(status))
        if rc != 0 and not (rc == _lib.MRBF_ENUMERIC and not raise_on_failure):
            if rc == _lib.MRBF_ENUMERIC and handle.value:
                self.lib.mrbf_free_model(self.ctx, handle)
            self._check(rc)
        return ModelBatch(self, handle.value), status

    def build_dev(self, cfg, sites, values, N, shape=None, status=None, recycle: Optional[ModelBatch] = None):
        import torch
        B, ts, n = sites.shape
        k = values.shape[2]
        if status is None:
            status = torch.zeros(B, dtype=torch.int32, device=sites.device)
        handle = C.c_void_p(recycle.handle if recycle is not None else None)
        if recycle is not None:
            recycle.handle = None
        ccfg = to_c_cfg(cfg)
        self._check(self.lib.mrbf_build_dev(self.ctx, C.byref(ccfg), B, n, k, ts, _ptr(N), _ptr(sites), _ptr(values), _ptr(shape),
                                            C.byref(handle), _ptr(status)))
        return ModelBatch(self, handle.value), status

    # ------------------------------------------------------------------ evaluation
    def eval(self, model: ModelBatch, X, want_values=True, want_jacobian=False):
        X = _np(X, np.float64).reshape(model.B, -1, model.n)
        M = X.shape[1]
        Y = np.zeros((model.B, M, model.k)) if want_values else None
        J = np.zeros((model.B, M, model.k, model.n)) if want_jacobian else None
        self._check(self.lib.mrbf_eval(self.ctx, model.handle, M, _ptr(X), _ptr(Y), _ptr(J)))
        return Y, J

    def eval_dev(self, model: ModelBatch, X, Y=None, J=None):
        M = X.shape[1]
        self._check(self.lib.mrbf_eval_dev(self.ctx, model.handle, M, _ptr(X), _ptr(Y), _ptr(J)))
        return Y, J

    # ------------------------------------------------------------------ steepest-descent direction (descent.jl:75-135)
    def descent_direction(self, jac, x, lb, ub, normalize: bool = True):
        """Batched _steepest_descent_direction: jac B x k x n, x B x n, lb / ub n -> (d B x n, omega B, iters B, status B)."""
        jac = _np(jac, np.float64)
        B, k, n = jac.shape
        x = _np(x, np.float64).reshape(B, n)
        lb = _np(np.broadcast_to(lb, (n,)), np.float64); ub = _np(np.broadcast_to(ub, (n,)), np.float64)
        d = np.zeros((B, n)); omega = np.zeros(B); iters = np.zeros(B, np.int32); status = np.zeros(B, np.int32)
        self._check(self.lib.mrbf_descent_direction(self.ctx, B, n, k, _ptr(jac), _ptr(x), _ptr(lb), _ptr(ub), int(bool(normalize)),
                                                    _ptr(d), _ptr(omega), _ptr(iters), _ptr(status)))
        return d, omega, iters, status

    def descent_direction_dev(self, jac, x, lb, ub, normalize: bool = True, out=None):
        """Device tensors in and out (enqueued on the engine's stream)."""
        import torch
        B, k, n = jac.shape
        if out is None:
            f64 = dict(dtype=torch.float64, device=jac.device); i32 = dict(dtype=torch.int32, device=jac.device)
            out = (torch.empty((B, n), **f64), torch.empty(B, **f64), torch.empty(B, **i32), torch.empty(B, **i32))
        self._check(self.lib.mrbf_descent_direction_dev(self.ctx, B, n, k, _ptr(jac), _ptr(x), _ptr(lb), _ptr(ub), int(bool(normalize)),
                                                        _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(out[3])))
        return out

    # ------------------------------------------------------------------ Pascoletti-Serafini inner solves
    def ps_solve(self, model: ModelBatch, x, lb, ub, mx=None, direction=None, n_obj: Optional[int] = None, objective: int = -1,
                 population: int = -1, max_evals: int = -1, seed: int = 0):
        """Batched `_ps_optimization` (direction given; descent.jl:478-500) or `_min_component` (direction None, one objective;
        descent.jl:369-387) on the surrogates: x, lb, ub B x n; mx, direction B x k.
        Returns (f_min B, x_min B x n, y_min B x k, found B, evals_per_instance)."""
        B, n, k = model.B, model.n, model.k
        x = _np(x, np.float64).reshape(B, n)
        lb = _np(np.broadcast_to(lb, (B, n)), np.float64); ub = _np(np.broadcast_to(ub, (B, n)), np.float64)
        mx_ = None if mx is None else _np(mx, np.float64).reshape(B, k)
        dir_ = None if direction is None else _np(np.broadcast_to(direction, (B, k)), np.float64)
        f = np.zeros(B); xm = np.zeros((B, n)); ym = np.zeros((B, k)); found = np.zeros(B, np.int32); used = C.c_int32(0)
        self._check(self.lib.mrbf_ps_solve(self.ctx, model.handle, _ptr(x), _ptr(lb), _ptr(ub), _ptr(mx_) if mx_ is not None else None,
                                           _ptr(dir_) if dir_ is not None else None, int(k if n_obj is None else n_obj), int(objective),
                                           int(population), int(max_evals), C.c_int64(int(seed)), _ptr(f), _ptr(xm), _ptr(ym), _ptr(found),
                                           C.byref(used)))
        return f, xm, ym, found, int(used.value)

    def ps_solve_dev(self, model: ModelBatch, x, lb, ub, mx=None, direction=None, n_obj: Optional[int] = None, objective: int = -1,
                     population: int = -1, max_evals: int = -1, seed: int = 0, out=None):
        """Device tensors in and out (enqueued on the engine's stream); returns (f_min, x_min, y_min, found) and the evaluation count."""
        import torch
        B, n, k = model.B, model.n, model.k
        if out is None:
            f64 = dict(dtype=torch.float64, device=x.device)
            out = (torch.empty(B, **f64), torch.empty((B, n), **f64), torch.empty((B, k), **f64), torch.empty(B, dtype=torch.int32, device=x.device))
        used = C.c_int32(0)
        self._check(self.lib.mrbf_ps_solve_dev(self.ctx, model.handle, _ptr(x), _ptr(lb), _ptr(ub), _ptr(mx) if mx is not None else None,
                                               _ptr(direction) if direction is not None else None, int(k if n_obj is None else n_obj),
                                               int(objective), int(population), int(max_evals), C.c_int64(int(seed)),
                                               _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(out[3]), C.byref(used)))
        return out, int(used.value)

    def backtrack_dev(self, model: ModelBatch, x, direction, step0, omega, armijo_c=1e-6, shrink=0.75,
                      min_stepsize=10 * np.finfo(np.float64).eps, max_loops=None, strict=True, out=None):
        """descent.jl:150-185 for device tensors (x, direction: B x n; step0, omega: B); returns
        (step_index, sigma, x_plus, mx, mx_plus) as device tensors, enqueued on the engine's stream."""
        import torch
        if max_loops is None:
            max_loops = int(math.floor(math.log(min_stepsize) / math.log(shrink)))
        B, n, k = model.B, model.n, model.k
        if out is None:
            f64 = dict(dtype=torch.float64, device=x.device)
            out = (torch.empty(B, dtype=torch.int32, device=x.device), torch.empty(B, **f64), torch.empty((B, n), **f64),
                   torch.empty((B, k), **f64), torch.empty((B, k), **f64))
        self._check(self.lib.mrbf_backtrack_dev(self.ctx, model.handle, _ptr(x), _ptr(direction), _ptr(step0), _ptr(omega),
                                                float(armijo_c), float(shrink), float(min_stepsize), int(max_loops), int(bool(strict)),
                                                _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(out[3]), _ptr(out[4])))
        return out

    # ------------------------------------------------------------------ device-resident database / model swap
    def db_append_dev(self, sites, values, n_db, new_sites, new_values, n_add, first_id=None, status=None):
        """new_result! for B device-resident databases (Databases.jl:174-183): appends the first n_add[b] rows of new_sites
        (B x add_stride x n) and new_values (B x add_stride x k, None => NaN = unevaluated); n_db is updated in place.
        Returns (first_id, status) -- status[b] = 1 when the capacity would be exceeded (nothing written)."""
        import torch
        B, db_stride, n = sites.shape
        k = values.shape[2]
        if first_id is None:
            first_id = torch.empty(B, dtype=torch.int32, device=sites.device)
        if status is None:
            status = torch.empty(B, dtype=torch.int32, device=sites.device)
        self._check(self.lib.mrbf_db_append_dev(self.ctx, B, n, k, db_stride, _ptr(sites), _ptr(values), _ptr(n_db), new_sites.shape[1],
                                                _ptr(new_sites), _ptr(new_values), _ptr(n_add), _ptr(first_id), _ptr(status)))
        return first_id, status

    def model_scatter_dev(self, dst: ModelBatch, src: ModelBatch, index_map, S: Optional[int] = None):
        """Instance s < S of `src` replaces instance index_map[s] of `dst` (negative: skipped), SurrogateContainer.jl:376-382."""
        self._check(self.lib.mrbf_model_scatter_dev(self.ctx, dst.handle, src.handle, _ptr(index_map),
                                                    int(index_map.shape[0] if S is None else S)))

    def backtrack(self, model: ModelBatch, x, direction, step0, omega, armijo_c=1e-6, shrink=0.75,
                  min_stepsize=10 * np.finfo(np.float64).eps, max_loops=None, strict=True):
        """descent.jl:150-185 with every step size evaluated in one launch."""
        if max_loops is None:
            max_loops = int(math.floor(math.log(min_stepsize) / math.log(shrink)))      # descent.jl:62-66
        B, n, k = model.B, model.n, model.k
        x = _np(x, np.float64).reshape(B, n); direction = _np(direction, np.float64).reshape(B, n)
        step0 = _np(np.broadcast_to(step0, (B,)), np.float64); omega = _np(np.broadcast_to(omega, (B,)), np.float64)
        idx = np.zeros(B, np.int32); sigma = np.zeros(B); xp = np.zeros((B, n)); mx = np.zeros((B, k)); mxp = np.zeros((B, k))
        self._check(self.lib.mrbf_backtrack(self.ctx, model.handle, _ptr(x), _ptr(direction), _ptr(step0), _ptr(omega),
                                            float(armijo_c), float(shrink), float(min_stepsize), int(max_loops), int(bool(strict)),
                                            _ptr(idx), _ptr(sigma), _ptr(xp), _ptr(mx), _ptr(mxp)))
        return xp, mxp, sigma[:, None] * direction, idx, mx
