"""Lock-step batched `optimize` over B independent multistart instances with a device-resident database
(BASELINE config C3: 4096 ZDT3 instances).

The reference runs one `optimize(mop, x0)` per instance, many of them concurrently under Threads.@threads
(examples/large_scale_benchmarks.jl:253).  Here every step of `iterate!` (src/algorithm.jl:615-917) is a batched call of
the C ABI over all instances, the databases (src/Databases.jl:15-32) live on the device for the whole run
(`mrbf_db_append_dev` = new_result!, the box scan runs inside the select kernels), and only the sites that need a true
function value cross to the host function and back.  What is vectorised host logic here is host logic in the
reference too (acceptance test, radius update, stopping tests, criticality loop): a few flops per instance.

Covered: one RbfConfig group holding all objectives, box constraints only (DummyFilter, no normal step), steepest
descent with Armijo backtracking (src/descent.jl:150-321), identity variable scaling (problem given in scaled space),
delta_max <= 1.  Everything else of `iterate!` (filter, restoration, Pascoletti-Serafini, var-scaler updates) stays
with the reference's host code.

  initialize_data / init_surrogates   src/algorithm.jl:223-313, src/SurrogateContainer.jl:272-295
  iterate!                            src/algorithm.jl:615-917
  criticality_routine                 src/algorithm.jl:523-612  (the shrunken radius is a local there: models are
                                      rebuilt with the iterate's unchanged radius, :572-579 -- reproduced)
  update_surrogates! / improve_...    src/SurrogateContainer.jl:334-391, src/models/RbfModel.jl:699-732
  AlgorithmConfig defaults            src/AbstractConfigInterface.jl:14-95 (Float32 literals promoted to Float64)

Instances move in lock-step through the phases; an instance that has stopped keeps its state and is masked out.
Sub-batches (model-improvement steps, criticality loops) are compacted, run through the same kernels, and their
models scattered back into the main model batch (`mrbf_model_scatter_dev`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Optional

import numpy as np

from .engine import Engine, ModelBatch, SelectResult, max_model_points

_f32 = lambda v: float(np.float32(v))
_SQRT_EPS32 = float(np.sqrt(np.float32(np.finfo(np.float32).eps)))
INT_MAX = 2**63 - 1

# return codes / iteration classifications (src/Morbit.jl enums; DB_FULL and NUMERIC are ours)
CONTINUE, MAX_ITER, BUDGET_EXHAUSTED, CRITICAL, TOLERANCE, INFEASIBLE, DB_FULL, NUMERIC = 0, 1, 2, 3, 4, 5, 6, 7
ACCEPTABLE, SUCCESSFULL, MODELIMPROVING, INACCEPTABLE, EARLY_EXIT, FILTER_ADD = 0, 1, 2, 3, 4, 5


@dataclass
class AlgorithmConfig:
    """src/ConfigImplementations.jl:13-85; defaults of src/AbstractConfigInterface.jl (Float32 literals)."""
    eps_crit: float = _f32(0.001)
    gamma_crit: float = _f32(0.51)
    max_critical_loops: int = 5
    delta_0: float = _f32(0.1)
    delta_max: float = _f32(0.5)
    max_evals: int = INT_MAX
    max_iter: int = 50
    f_tol_rel: float = _SQRT_EPS32
    x_tol_rel: float = _SQRT_EPS32
    f_tol_abs: float = -1.0
    x_tol_abs: float = -1.0
    omega_tol_rel: float = float(np.float32(10) * np.sqrt(np.float32(np.finfo(np.float32).eps)))
    delta_tol_rel: float = _SQRT_EPS32
    omega_tol_abs: float = -math.inf
    delta_tol_abs: float = _SQRT_EPS32
    stepnorm_tol_abs: float = 0.0
    strict_acceptance_test: bool = True
    nu_success: float = _f32(0.2)
    nu_accept: float = 0.0
    mu: float = _f32(2e3)
    beta: float = _f32(1e3)
    gamma_grow: float = 2.0
    gamma_shrink: float = 0.75
    gamma_shrink_much: float = _f32(0.51)
    # SteepestDescentConfig, src/descent.jl:51-73
    strict_backtracking: bool = True
    armijo_const_rhs: float = 1e-6
    armijo_const_shrink: float = 0.75
    min_stepsize: float = 10 * float(np.finfo(np.float64).eps)
    normalize: bool = True
    # rtol of `Δ ≈ Δ_max` (RbfModel.jl:588): these defaults are the reference's Float32 literals, so Julia compares a Float64
    # radius with a Float32 delta_max => sqrt(eps(Float32)); 1.4901161193847656e-08 for an AlgorithmConfig{Float64}
    isapprox_rtol: float = float(np.sqrt(np.float32(np.finfo(np.float32).eps)))


def _intersect_box_absmax(torch, x, d, lb, ub):
    """intersect_box(x, d, lb, ub; return_vals = :absmax) for a batch (utilities.jl:126-221, 285-287): rows of x, d, lb, ub."""
    inf = math.inf
    nz = d != 0
    tl, tu = lb - x, ub - x
    sl = torch.where(tl == 0, torch.where(d > 0, inf, 0.0), tl / d)
    su = torch.where(tu == 0, torch.where(d < 0, inf, 0.0), tu / d)
    sig = torch.cat([sl, su], 1)
    valid = torch.cat([nz, nz], 1)
    pos = valid & (sig >= 0)
    neg = valid & ~(sig >= 0)
    s_pos = torch.where(pos, sig, inf).min(1).values
    s_pos = torch.where(pos.any(1), s_pos, 0.0)
    s_neg = torch.where(neg, sig, -inf).max(1).values
    s_neg = torch.where(neg.any(1), s_neg, 0.0)
    out = torch.where(s_pos.abs() >= s_neg.abs(), s_pos, s_neg)          # positive wins ties, :212-217
    return torch.where(nz.any(1), out, inf)


class _Phase:
    """Wall-clock accounting of the driver's phases (only when LockstepDriver(profile=True): it synchronises around every phase)."""
    def __init__(self, drv, name):
        self.drv, self.name = drv, name

    def __enter__(self):
        if self.drv.profile:
            import time
            self.drv.torch.cuda.synchronize(self.drv.dev)
            self.t0 = time.perf_counter()

    def __exit__(self, *exc):
        if self.drv.profile:
            import time
            self.drv.torch.cuda.synchronize(self.drv.dev)
            self.drv.phase_s[self.name] = self.drv.phase_s.get(self.name, 0.0) + time.perf_counter() - self.t0
        return False


class _Scratch:
    """Per-batch-size scratch: select outputs, kept factorisation, model batch (reused: no allocation per iteration)."""
    def __init__(self):
        self.sel: Optional[SelectResult] = None
        self.prepared = None
        self.model: Optional[ModelBatch] = None
        self.train = None
        self.model_scratch: Optional[ModelBatch] = None


class LockstepDriver:
    """B independent `optimize` runs in lock-step on one GPU.

    func      maps an (M, n) array of scaled sites to (M, k) values: NumPy in / NumPy out (the reference's host boundary,
              eval_missing!, src/Databases.jl:258-277) or, with `func_on_device=True`, torch CUDA tensors in and out.
    x0        (B, n) starting points (projected into the box like algorithm.jl:256-258).
    capacity  database rows per instance (fixed; an instance that would outgrow it stops with DB_FULL).
    """

    def __init__(self, cfg, func: Callable, x0, glb, gub, ac: Optional[AlgorithmConfig] = None, device: str = "cuda:0",
                 capacity: int = 128, func_on_device: bool = False, record: bool = False, engine: Optional[Engine] = None,
                 profile: bool = False, scratch: Optional[dict] = None):
        import torch
        self.torch = torch
        if engine is None:
            # the C ABI's kernels and the driver's torch bookkeeping are ordered on ONE stream: torch's current stream
            dev_ = torch.device(device)
            engine = Engine(dev_.index or 0, stream=torch.cuda.current_stream(dev_).cuda_stream)
        self.engine, self.cfg, self.func, self.ac = engine, cfg, func, ac or AlgorithmConfig()
        self.engine.set_isapprox_rtol(self.ac.isapprox_rtol)
        self.func_on_device, self.record = func_on_device, record
        self.n_func_calls = 0
        self.profile, self.phase_s = profile, {}
        self.n_sites_evaluated = 0
        if self.ac.delta_max > 1.0:
            raise ValueError("delta_max > 1: compute_descent_step's second branch (descent.jl:276-311) is not covered")
        dev = torch.device(device)
        self.dev = dev
        f64, i32 = dict(dtype=torch.float64, device=dev), dict(dtype=torch.int32, device=dev)
        x0 = np.asarray(x0, np.float64)
        B, n = x0.shape
        self.B, self.n, self.cap = B, n, int(capacity)
        self.glb = torch.from_numpy(np.array(np.broadcast_to(np.asarray(glb, np.float64), (n,)))).to(dev)
        self.gub = torch.from_numpy(np.array(np.broadcast_to(np.asarray(gub, np.float64), (n,)))).to(dev)
        self.x = torch.minimum(torch.maximum(torch.from_numpy(x0).to(dev), self.glb), self.gub).contiguous()
        self.fx = self._f(self.x)
        k = self.fx.shape[1]
        self.k = k
        self.num_evals = torch.ones(B, dtype=torch.int64, device=dev)
        self.sites = torch.zeros((B, self.cap, n), **f64)
        self.values = torch.full((B, self.cap, k), math.nan, **f64)
        self.n_db = torch.zeros(B, **i32)
        self.x_index = torch.ones(B, **i32)
        self.delta = torch.full((B,), float(self.ac.delta_0), **f64)
        self.ret = torch.zeros(B, **i32)
        self.it_stat = torch.zeros(B, **i32)
        self.iter_counter = 1
        self.iters_done = torch.zeros(B, **i32)
        mp = max(1, min(max_model_points(cfg, n), self.cap))     # same width as the select outputs (engine.select_points_keep_dev)
        # RbfMeta of every instance (src/models/RbfModel.jl:148-159)
        self.r1 = torch.zeros((B, n), **i32); self.n_r1 = torch.zeros(B, **i32)
        self.r2 = torch.zeros((B, n), **i32); self.n_r2 = torch.zeros(B, **i32)
        self.r3_sites = torch.zeros((B, n, n), **f64); self.r3_values = torch.zeros((B, n, k), **f64); self.n_r3 = torch.zeros(B, **i32)
        self.r3_first = torch.zeros(B, **i32); self.center = torch.ones(B, **i32)
        self.r4 = torch.zeros((B, mp), **i32); self.n_r4 = torch.zeros(B, **i32)
        self.dirs = torch.zeros((B, n, n), **f64); self.n_dirs = torch.zeros(B, **i32)
        self.fully_linear = torch.zeros(B, dtype=torch.bool, device=dev)
        self.build_failures = torch.zeros(B, **i32)
        self.model: Optional[ModelBatch] = None
        self._scratch = scratch if scratch is not None else {}      # per-size device scratch; may be handed on from an earlier driver
        self._all = torch.arange(B, device=dev)
        self._first_id = torch.zeros(B, **i32); self._app_status = torch.zeros(B, **i32)
        self._add_sites = torch.zeros((B, n, n), **f64); self._add_values = torch.zeros((B, n, k), **f64); self._n_add = torch.zeros(B, **i32)
        self._J = torch.zeros((B, 1, k, n), **f64)
        self._lp_out = None
        self._bt_out = None
        self._X2 = torch.zeros((B, 2, n), **f64); self._Y2 = torch.zeros((B, 2, k), **f64)
        self.numeric_log: List[dict] = []
        self.trace: List[dict] = []          # record=True: host copies of the state after every iterate()
        self.lp_calls: List[dict] = []       # record=True: every criticality computation in order (mask, d, omega)
        # initialize_data: x0 is result #1 of every database (build_super_db, utilities.jl:39)
        self._append(self.x[:, None, :], self.fx[:, None, :], torch.ones(B, **i32))
        self._update(None, True)             # init_surrogates: prepare_init_model => ensure_fully_linear = true (RbfModel.jl:506-513)

    # ------------------------------------------------------------------ host <-> user function
    def _f(self, X):
        """Values of the true objectives at the rows of X (device tensor) -> (M, k) device tensor."""
        torch = self.torch
        self.n_func_calls += 1
        self.n_sites_evaluated += int(X.shape[0])
        with _Phase(self, "objective_function"):
            if self.func_on_device:
                return self.func(X).to(torch.float64).contiguous()
            Y = np.asarray(self.func(X.cpu().numpy()), np.float64)
            return torch.from_numpy(np.ascontiguousarray(Y.reshape(X.shape[0], -1))).to(self.dev)

    # ------------------------------------------------------------------ database
    def _append(self, new_sites, new_values, n_add):
        """new_result! for every instance (Databases.jl:174-183); returns the id of the first appended row."""
        self.engine.db_append_dev(self.sites, self.values, self.n_db, new_sites.contiguous(), None if new_values is None else new_values.contiguous(),
                                  n_add, self._first_id, self._app_status)
        full = self._app_status != 0
        self.ret = self.torch.where(full & (self.ret == CONTINUE), DB_FULL, self.ret)
        return self._first_id.clone()

    # ------------------------------------------------------------------ surrogates
    def _subset(self, idx):
        """Pads an index list to a power of two (scratch buffers are cached per size); returns (S, padded index, map)."""
        torch = self.torch
        S = int(idx.numel())
        Sp = 16
        while Sp < S:
            Sp *= 2
        Sp = min(Sp, max(self.B, 16))
        if Sp > S:
            pad = idx[:1].expand(Sp - S)
            idx_p = torch.cat([idx, pad])
            mp_ = torch.cat([idx, torch.full((Sp - S,), -1, device=self.dev, dtype=idx.dtype)])
        else:
            idx_p, mp_ = idx, idx
        return S, idx_p, mp_.to(torch.int32)

    def _max_new(self, idx_p):
        lim = min(self.ac.max_evals, self.cfg.max_evals)
        left = (lim - 1) - self.num_evals[idx_p]                          # RbfModel.jl:613-618 (nothing is unevaluated here)
        return left.clamp(0, 2**31 - 1).to(self.torch.int32)

    def _update(self, idx, ensure_fully_linear: bool):
        """update_surrogates! (SurrogateContainer.jl:339-390) for the instances `idx` (None = all): rounds 1-4 with the kept
        factorisation, true values of the new round-3 sites, coefficient solve; models scattered into the main batch."""
        torch, E = self.torch, self.engine
        if idx is None:
            S, idx_p, imap = self.B, self._all, None
            sites, values, n_db, x_index, x, delta = self.sites, self.values, self.n_db, self.x_index, self.x, self.delta
        else:
            if idx.numel() == 0:
                return
            S, idx_p, imap = self._subset(idx)
            sites, values = self.sites.index_select(0, idx_p), self.values.index_select(0, idx_p)
            n_db, x_index = self.n_db.index_select(0, idx_p), self.x_index.index_select(0, idx_p)
            x, delta = self.x.index_select(0, idx_p), self.delta.index_select(0, idx_p)
        Sp = int(idx_p.numel())
        sc = self._scratch.setdefault((Sp, self.n, self.k, self.cap, self.cfg.kernel, self.cfg.max_model_points, self.cfg.polynomial_degree), _Scratch())
        flags = torch.zeros((Sp, 2), dtype=torch.int32, device=self.dev)
        flags[:, 0] = 1 if ensure_fully_linear else 0
        with _Phase(self, "select_rounds_1_4"):
            sel, sc.prepared = E.select_points_keep_dev(self.cfg, sites, n_db, x_index, x, delta, float(self.ac.delta_max), self.glb,
                                                        self.gub, flags, self._max_new(idx_p), out=sc.sel, prepared=sc.prepared)
        sc.sel = sel
        n = self.n
        new = torch.arange(n, device=self.dev)[None, :] < sel.n_r3[:, None]            # (Sp, n) rows of r3_sites that are new sites
        if idx is not None:
            new[S:] = False
        r3_values = torch.zeros((Sp, n, self.k), dtype=torch.float64, device=self.dev)
        if bool(new.any()):
            r3_values[new] = self._f(sel.r3_sites[new])                                 # eval_missing!, Databases.jl:258-277
        status = torch.zeros(Sp, dtype=torch.int32, device=self.dev)
        main_missing = self.model is None
        with _Phase(self, "build_from_kept_factor"):
            mdl, status = E.build_prepared_dev(self.cfg, sc.prepared, sites, values, x_index, sel, r3_values, status,
                                               recycle=sc.model)
        bad = (sel.status != 0) | (status != 0)
        if self.record and bool(bad[:S].any()):
            self.numeric_log.append(dict(iter=self.iter_counter, select=int((sel.status[:S] != 0).sum()), build=int((status[:S] != 0).sum()),
                                         build_codes=status[:S][status[:S] != 0][:8].cpu().tolist()))
        # new_result! of the round-3 sites (ids n_db+1.., RbfModel.jl:301-305), now with their values
        tgt = idx_p[:S]
        self._n_add.zero_()
        self._n_add[tgt] = sel.n_r3[:S]
        self._add_sites[tgt] = sel.r3_sites[:S]
        self._add_values[tgt] = r3_values[:S]
        first = self._append(self._add_sites, self._add_values, self._n_add)
        self.num_evals[tgt] += sel.n_r3[:S].to(torch.int64)
        # commit the meta data -- only where the coefficient solve succeeded.  A failed solve (reduced kernel matrix not positive
        # definite: the reference's own training sets become numerically singular on exactly structured databases, its LU then
        # returns coefficients with cond ~ 1e18) keeps the instance's previous model and meta data; the instance goes on and the
        # event is counted in `build_failures`.  Without a previous model (initialisation) the instance stops with NUMERIC.
        ok = ~bad[:S]
        if main_missing:
            self.ret[tgt] = torch.where(bad[:S] & (self.ret[tgt] == CONTINUE), NUMERIC, self.ret[tgt])
        self.build_failures[tgt] += bad[:S].to(torch.int32)
        tk = tgt[ok]
        for name in ("r1", "n_r1", "r2", "n_r2", "n_r3", "r4", "n_r4", "dirs", "n_dirs", "r3_sites"):
            getattr(self, name)[tk] = getattr(sel, name)[:S][ok]
        self.r3_values[tk] = r3_values[:S][ok]
        self.r3_first[tk] = first[tk]
        self.center[tk] = x_index[:S][ok]
        self.fully_linear[tk] = sel.flags_out[:S, 0][ok] != 0
        self.fully_linear[tgt[~ok]] = False
        if main_missing:
            # main model batch: room for the training set of the select kernels plus the sites that model-improvement steps
            # push onto round1_indices afterwards (RbfModel.jl:719); allocated once through a placeholder build
            ts = mdl.train_stride + n
            z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=self.dev)
            self.model, _ = E.build_dev(self.cfg, z(self.B, ts, n), z(self.B, ts, self.k), torch.ones(self.B, dtype=torch.int32, device=self.dev))
        sc.model = mdl
        if imap is None:
            imap = self._all.to(torch.int32)
        imap = imap.clone()
        imap[:S] = torch.where(ok, imap[:S], -1)
        E.model_scatter_dev(self.model, mdl, imap.contiguous(), Sp)

    def _improve(self, idx):
        """improve_surrogates!: prepare_improve_model (RbfModel.jl:699-732) + update_model from the meta's index lists."""
        torch, E, cfg = self.torch, self.engine, self.cfg
        if idx.numel() == 0:
            return
        n, k = self.n, self.k
        x, delta = self.x[idx], self.delta[idx]
        do = ~self.fully_linear[idx] & (self.n_dirs[idx] > 0)
        delta_1 = delta * cfg.theta_enlarge_1
        lb_1 = torch.maximum(self.glb[None, :], x - delta_1[:, None]); ub_1 = torch.minimum(self.gub[None, :], x + delta_1[:, None])
        piv = delta_1 * cfg.theta_pivot
        dirs = self.dirs[idx]
        d = dirs[:, 0, :]
        length = _intersect_box_absmax(torch, x, d, lb_1, ub_1)
        offset = length[:, None] * d
        n_train = 1 + self.n_r1[idx] + self.n_r2[idx] + self.n_r3[idx] + self.n_r4[idx]
        success = do & (offset.abs().max(1).values > piv) & (n_train < self.model.train_stride) & (self.n_db[idx] < self.cap)
        # popfirst!(meta.improving_directions)
        popped = torch.cat([dirs[:, 1:, :], torch.zeros_like(dirs[:, :1, :])], 1)
        self.dirs[idx] = torch.where(do[:, None, None], popped, dirs)
        self.n_dirs[idx] = self.n_dirs[idx] - do.to(torch.int32)
        if bool(success.any()):
            new_site = (x + offset)
            ok = success.nonzero().flatten()
            vals = self._f(new_site[ok])
            self._n_add.zero_()
            tgt = idx[ok]
            self._n_add[tgt] = 1
            self._add_sites[tgt, 0] = new_site[ok]
            self._add_values[tgt, 0] = vals
            first = self._append(self._add_sites, self._add_values, self._n_add)
            pos = self.n_r1[tgt].to(torch.int64)
            self.r1[tgt, pos] = first[tgt]                                            # push!(meta.round1_indices, new_id)
            self.n_r1[tgt] += 1
            self.num_evals[tgt] += 1
            self.fully_linear[tgt] = self.fully_linear[tgt] | (self.n_dirs[tgt] == 0)
            # `meta.improving_directions` still holds the directions round 3 has used (RbfModel.jl:575-577, 625), so the new site can
            # be an exact copy of a round-3 site: the interpolation system is then singular by construction.  That solve is not
            # attempted; it counts as a failed build (previous model kept, see _update)
            valid = torch.arange(n, device=self.dev)[None, :] < self.n_r3[tgt][:, None]
            dup = ((self.r3_sites[tgt] == new_site[ok][:, None, :]).all(-1) & valid).any(-1)
            self.build_failures[tgt[dup]] += 1
            self.fully_linear[tgt[dup]] = False
            success = success.clone()
            success[ok[dup]] = False
        # improve_model = update_model: from-scratch solve of [centre; r1; r2; r3; r4].  Only the instances whose training set changed
        # are solved again: for the others update_model would reproduce the model they already hold
        if not bool(success.any()):
            return
        idx = idx[success]
        S, idx_p, imap = self._subset(idx)
        Sp = int(idx_p.numel())
        sc = self._scratch.setdefault((Sp, self.n, self.k, self.cap, self.cfg.kernel, self.cfg.max_model_points, self.cfg.polynomial_degree), _Scratch())
        g = lambda t: t.index_select(0, idx_p)
        sel = SelectResult(g(self.r1), g(self.n_r1), g(self.r2), g(self.n_r2), g(self.r3_sites), g(self.n_r3), g(self.r4), g(self.n_r4),
                           None, None, None, None)
        ts = self.model.train_stride
        if sc.train is not None and sc.train[0].shape != (Sp, ts, n):
            sc.train = None
        sc.train = E.gather_training_dev(g(self.sites), g(self.values), g(self.x_index), sel, g(self.r3_values), ts, out=sc.train)
        status = torch.zeros(Sp, dtype=torch.int32, device=self.dev)
        with _Phase(self, "improve_build_from_scratch"):
            sc.model_scratch, status = E.build_dev(cfg, sc.train[0], sc.train[1], sc.train[2], None, status, recycle=sc.model_scratch)
        ok = status[:S] == 0                               # a failed solve keeps the previous model (see _update)
        self.build_failures[idx] += (~ok).to(torch.int32)
        self.fully_linear[idx[~ok]] = False
        imap = imap.clone()
        imap[:S] = torch.where(ok, imap[:S], -1)
        E.model_scatter_dev(self.model, sc.model_scratch, imap.contiguous(), Sp)

    # ------------------------------------------------------------------ criticality (descent.jl:187-241)
    def _criticality(self, mask=None, rec=True):
        """Jacobian at the iterate + exact LP for every instance; (omega, d).  `mask` only labels the recorded call."""
        E = self.engine
        with _Phase(self, "jacobian_lp"):
            E.eval_dev(self.model, self.x[:, None, :].contiguous(), None, self._J)
            self._lp_out = E.descent_direction_dev(self._J.view(self.B, self.k, self.n), self.x, self.glb, self.gub, self.ac.normalize,
                                                   out=self._lp_out)
        d, omega = self._lp_out[0].clone(), self._lp_out[1].clone()
        if self.record and rec:
            m = (self.ret == CONTINUE) if mask is None else mask
            self.lp_calls.append(dict(mask=m.cpu().numpy().copy(), d=d.cpu().numpy(), omega=omega.cpu().numpy()))
        return omega, d

    def _budget_okay(self):
        return self.num_evals < min(self.cfg.max_evals, self.ac.max_evals)

    def _criticality_routine(self, C, omega, d):
        """criticality_routine (algorithm.jl:523-612) for the instances in mask C; returns (exit mask, omega, d, loops)."""
        torch, ac = self.torch, self.ac
        beta = max(ac.beta, ac.mu)
        loops = torch.zeros(self.B, dtype=torch.int32, device=self.dev)
        exit_c = torch.zeros_like(C)
        do_loops = C.clone()
        nfl = C & ~self.fully_linear
        self._fail_mark = self.build_failures.clone()
        if bool(nfl.any()):
            self._update(nfl.nonzero().flatten(), True)
            om2, d2 = self._criticality(nfl)
            omega = torch.where(nfl, om2, omega); d = torch.where(nfl[:, None], d2, d)
            do_loops = torch.where(nfl, self.fully_linear & (self.delta > ac.mu * omega), do_loops)
        delta = self.delta.clone()
        delta_0 = self.delta.clone()
        in_loop = do_loops & (delta > ac.mu * omega) & (self.ret == CONTINUE)
        # The reference rebuilds the models in every loop with the iterate's UNCHANGED radius (:572-579), i.e. with the same inputs
        # as the loop before unless that update appended new round-3 sites.  An update that appended nothing is a fixed point
        # (same database, iterate, radius, flags => same training set, model, omega), so it is not recomputed -- exact, not a
        # heuristic; the loop counters, radius and exit tests still advance as in the reference.
        fixed = nfl & (self.n_r3 == 0) & (self.build_failures == self._fail_mark)
        while bool(in_loop.any()):
            stop = in_loop & ((loops >= ac.max_critical_loops) | ~self._budget_okay())
            exit_c |= stop
            in_loop &= ~stop
            if not bool(in_loop.any()):
                break
            delta = torch.where(in_loop, ac.gamma_crit * delta, delta)
            need = in_loop & ~fixed
            if bool(need.any()):
                self._fail_mark = self.build_failures.clone()
                self._update(need.nonzero().flatten(), True)      # the iterate keeps its radius (reference quirk, :572-579)
                om2, d2 = self._criticality(need, rec=False)
                omega = torch.where(need, om2, omega); d = torch.where(need[:, None], d2, d)
                fixed = fixed | (need & (self.n_r3 == 0) & (self.build_failures == self._fail_mark))
            if self.record:
                self.lp_calls.append(dict(mask=in_loop.cpu().numpy().copy(), d=d.cpu().numpy(), omega=omega.cpu().numpy()))
            loops += in_loop.to(torch.int32)
            tol = in_loop & ((delta <= ac.delta_tol_abs) | ((omega <= ac.omega_tol_rel) & (delta <= ac.delta_tol_rel))
                             | (omega <= ac.omega_tol_abs) | ~self.fully_linear)
            exit_c |= tol
            in_loop &= ~tol
            in_loop &= (delta > ac.mu * omega) & (self.ret == CONTINUE)
        self.delta = torch.where(do_loops, torch.minimum(delta_0, torch.maximum(beta * omega, delta)), self.delta)   # :603
        return exit_c, omega, d, loops

    # ------------------------------------------------------------------ iterate! (algorithm.jl:615-917)
    def iterate(self):
        torch, ac, E = self.torch, self.ac, self.engine
        B = self.B
        W = lambda m, a, b: torch.where(m, a, b)
        active = self.ret == CONTINUE
        if not bool(active.any()):
            return False
        self.iters_done += active.to(torch.int32)
        stat = self.it_stat.clone()
        nan = torch.full((B,), math.nan, dtype=torch.float64, device=self.dev)
        rho, steplength, omega_rec = nan.clone(), nan.clone(), nan.clone()

        def stop(mask, code):
            nonlocal active, stat
            m = mask & active & (self.ret == CONTINUE)
            self.ret = W(m, code, self.ret)
            stat = W(m, EARLY_EXIT, stat)
            active = active & ~m

        stop(torch.full_like(active, self.iter_counter > ac.max_iter), MAX_ITER)
        stop(~self._budget_okay(), BUDGET_EXHAUSTED)
        stop(self.delta <= ac.delta_tol_abs, TOLERANCE)
        stop(self.n_db + self.n + 2 > self.cap, DB_FULL)            # room for a round-3 rebuild (n sites) + the trial point
        beta = max(ac.beta, ac.mu)
        if bool(active.any()):
            if self.iter_counter > 1:
                imp = active & (self.it_stat == MODELIMPROVING)
                upd = active & ~imp
                if bool(upd.all()):
                    self._update(None, False)
                else:
                    self._update(upd.nonzero().flatten(), False)
                self._improve(imp.nonzero().flatten())
                active = active & (self.ret == CONTINUE)
            omega, d = self._criticality(active)
            omega_rec = W(active, omega, omega_rec)
            stop(((omega <= ac.omega_tol_rel) & (self.delta <= ac.delta_tol_rel)) | (omega <= ac.omega_tol_abs), CRITICAL)
            loops = torch.zeros(B, dtype=torch.int32, device=self.dev)
            C = active & (omega <= ac.eps_crit) & (~self.fully_linear | (self.delta > ac.mu * omega))
            if bool(C.any()):
                exit_c, omega, d, loops = self._criticality_routine(C, omega, d)
                omega_rec = W(C, omega, omega_rec)
                stop(exit_c, CRITICAL)
                active = active & (self.ret == CONTINUE)
        if bool(active.any()):
            # compute_descent_step, descent.jl:243-321 (delta <= 1 branch), then _backtrack with all step sizes in one launch
            x = self.x
            norm_d = d.abs().max(1).values
            sigma = W(norm_d > 0, torch.minimum(self.delta / norm_d, torch.ones_like(norm_d)), torch.ones_like(norm_d))
            small = ~(sigma > ac.min_stepsize)
            with _Phase(self, "backtrack"):
                self._bt_out = E.backtrack_dev(self.model, x, d.contiguous(), sigma.contiguous(), omega.contiguous(), ac.armijo_const_rhs,
                                               ac.armijo_const_shrink, ac.min_stepsize, None, ac.strict_backtracking, out=self._bt_out)
            x_trial = W(small[:, None], x, self._bt_out[2])
            omega = W(small, torch.zeros_like(omega), omega)
            act_idx = active.nonzero().flatten()
            fx_trial = self.fx.clone()
            fx_trial[act_idx] = self._f(x_trial[act_idx])                                # :760
            self.num_evals += active.to(torch.int64)
            self._n_add.zero_(); self._n_add[act_idx] = 1
            self._add_sites[:, 0] = x_trial; self._add_values[:, 0] = fx_trial
            new_index = self._append(self._add_sites, self._add_values, self._n_add)     # put_eval_result_into_db!, :764
            active = active & (self.ret == CONTINUE)
            self._X2[:, 0] = x; self._X2[:, 1] = x_trial
            with _Phase(self, "rho_evals"):
                E.eval_dev(self.model, self._X2, self._Y2, None)                             # :766-767
            mx, mx_trial = self._Y2[:, 0], self._Y2[:, 1]
            steplength = W(active, (x - x_trial).abs().max(1).values, steplength)        # :773
            if ac.strict_acceptance_test:
                denom = mx - mx_trial
                r_ = ((self.fx - fx_trial) / denom).min(1).values
                r_ = W((denom == 0).any(1), nan, r_)
                good = (denom >= 0).all(1)
            else:
                denom = mx.max(1).values - mx_trial.max(1).values
                r_ = (self.fx.max(1).values - fx_trial.max(1).values) / denom
                good = denom >= 0
            r_ = W(torch.isnan(r_), torch.full_like(r_, -math.inf), r_)
            rho = W(active, r_, rho)
            succ = r_ >= ac.nu_success
            fl = self.fully_linear
            acc_ok = r_ >= ac.nu_accept
            # classification, :812-866 (DummyFilter: always acceptable, theta_k = 0)
            new_stat = W(good, W(succ, SUCCESSFULL, W(fl, W(acc_ok, ACCEPTABLE, INACCEPTABLE), MODELIMPROVING)), FILTER_ADD).to(torch.int32)
            accept = W(good, succ | (fl & acc_ok), torch.ones_like(good))
            grow = (good & succ & (self.delta < beta * omega)) | (~good & succ)
            shrink = good & ~succ & fl & acc_ok
            shrink_much = good & ~succ & fl & ~acc_ok
            stat = W(active, new_stat, stat)
            tol_exit = active & ~accept & (steplength <= ac.stepnorm_tol_abs)            # :872-876
            self.ret = W(tol_exit, TOLERANCE, self.ret)
            go = active & ~tol_exit
            new_delta = W(grow, torch.clamp(ac.gamma_grow * self.delta, max=float(ac.delta_max)),
                          W(shrink, self.delta * ac.gamma_shrink, W(shrink_much, self.delta * ac.gamma_shrink_much, self.delta)))
            self.delta = W(go, new_delta, self.delta)
            acc = go & accept
            dx = (x - x_trial).abs().max(1).values
            df = (self.fx - fx_trial).abs().max(1).values
            tol2 = acc & ((dx <= ac.x_tol_rel * x.abs().max(1).values) | (dx <= ac.x_tol_abs)
                          | (df <= ac.f_tol_rel * self.fx.abs().max(1).values) | (df <= ac.f_tol_abs))      # :905-912
            self.x = W(acc[:, None], x_trial, x).contiguous()
            self.fx = W(acc[:, None], fx_trial, self.fx)
            self.x_index = W(acc, new_index, self.x_index)
            self.ret = W(tol2, TOLERANCE, self.ret)
        self.it_stat = stat.to(torch.int32)
        if self.record:
            c = lambda t: t.detach().cpu().numpy().copy()
            self.trace.append(dict(iter_counter=self.iter_counter, ret=c(self.ret), it_stat=c(self.it_stat), x=c(self.x), fx=c(self.fx),
                                   x_index=c(self.x_index), delta=c(self.delta), n_db=c(self.n_db), num_evals=c(self.num_evals),
                                   omega=c(omega_rec), rho=c(rho), steplength=c(steplength), fully_linear=c(self.fully_linear),
                                   center=c(self.center), r1=c(self.r1), n_r1=c(self.n_r1), r2=c(self.r2), n_r2=c(self.n_r2),
                                   r3_first=c(self.r3_first), n_r3=c(self.n_r3), r4=c(self.r4), n_r4=c(self.n_r4)))
        self.iter_counter += 1
        return bool((self.ret == CONTINUE).any())

    def run(self, max_steps: Optional[int] = None):
        """Iterates until every instance has stopped (or max_steps lock-step iterations); returns (x, fx, ret_code) as NumPy."""
        steps = 0
        while self.iterate():
            steps += 1
            if max_steps is not None and steps >= max_steps:
                break
        return self.x.cpu().numpy(), self.fx.cpu().numpy(), self.ret.cpu().numpy()

    def training_ids(self, b: int) -> List[int]:
        """_collect_indices(meta) of instance b (RbfModel.jl:178-186); round-3 sites by their database ids."""
        c = lambda t: t[b].cpu().numpy()
        r3 = [int(self.r3_first[b]) + i for i in range(int(self.n_r3[b]))]
        return ([int(self.center[b])] + [int(v) for v in c(self.r1)[: int(self.n_r1[b])]] + [int(v) for v in c(self.r2)[: int(self.n_r2[b])]]
                + r3 + [int(v) for v in c(self.r4)[: int(self.n_r4[b])]])
