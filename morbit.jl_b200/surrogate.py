"""Host-side mirror of Morbit's surrogate plugin interface for the GPU RbfConfig.

Same names, argument meaning and error behaviour as the reference methods a
`GpuRbfConfig <: AbstractSurrogateConfig` has to implement (src/AbstractSurrogateInterface.jl:6-79;
reference implementations in src/models/RbfModel.jl).  The reference's host stays Julia (julia/GpuRbf.jl
shows the `ccall` shim); Julia is not available in this image, so this Python mirror is what the parity
tests drive.  Everything numerical goes through the C ABI in libmorbit_rbf.so; only the database
bookkeeping (ids, unevaluated list, site matching) is host logic, exactly as in the Julia shim.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .engine import Engine, ModelBatch

INT_MAX = 2**63 - 1
RBF_KERNELS = ("cubic", "inv_multiquadric", "multiquadric", "thin_plate_spline", "gaussian")   # RbfModel.jl:48-54

_default_engine: Optional[Engine] = None


def default_engine() -> Engine:
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(0)
    return _default_engine


# ---------------------------------------------------------------------------------------- config / meta
@dataclass(frozen=True)
class RbfConfig:
    """RbfConfig, src/models/RbfModel.jl:66-112 (same fields, defaults and asserts)."""
    kernel: str = "cubic"
    shape_parameter: object = float("nan")          # Float64 or a String in Δ (RbfModel.jl:135-143)
    polynomial_degree: int = 1
    theta_enlarge_1: float = 2.0
    theta_enlarge_2: float = 2.0
    theta_pivot: Optional[float] = None
    theta_pivot_cholesky: float = 1e-7
    require_linear: bool = True
    max_model_points: int = -1
    use_max_points: bool = False
    optimized_sampling: bool = True
    max_evals: int = INT_MAX

    def __post_init__(self):
        if self.theta_pivot is None:
            object.__setattr__(self, "theta_pivot", 1.0 / (2.0 * self.theta_enlarge_1))
        sp = self.shape_parameter
        assert self.theta_enlarge_1 * self.theta_pivot <= 1, "θ_pivot must be <= θ_enlarge_1^(-1)."
        assert self.kernel in RBF_KERNELS, "`kernel` not supported. See `RBF_KERNELS` for available symbols."
        if not isinstance(sp, str):
            nan = math.isnan(sp)
            if self.kernel == "thin_plate_spline":
                assert nan or (sp % 1 == 0 and sp >= 1), "Invalid shape_parameter for :thin_plate_spline."
            if self.kernel == "cubic":
                assert nan or (sp % 1 == 0 and sp % 2 == 1), "Invalid shape_parameter for :cubic."
            assert nan or sp > 0, "Shape parameter must be strictly positive."
        assert self.theta_enlarge_1 >= 1 and self.theta_enlarge_2 >= 1, "θ's must be >= 1."

    def signature(self):                             # _get_signature, :114
        return (self.theta_pivot, self.theta_enlarge_1, self.theta_enlarge_2, self.optimized_sampling)


def max_evals(cfg: RbfConfig) -> int:               # :120
    return cfg.max_evals


def combinable(cfg: RbfConfig) -> bool:             # :121
    return True


def parse_shape_param_string(delta: float, expr: str) -> float:
    """RbfModel.jl:135-143: evaluate a string such as "10/Δ" with Δ bound to the radius."""
    return float(eval(expr.replace("Δ", "Delta").replace("^", "**"), {"__builtins__": {}}, {"Delta": delta, "sqrt": math.sqrt}))


def _shape_value(delta: float, cfg: RbfConfig) -> float:
    sp = cfg.shape_parameter
    return parse_shape_param_string(delta, sp) if isinstance(sp, str) else float(sp)


@dataclass
class RbfMeta:
    """RbfMeta, RbfModel.jl:148-159."""
    signature: tuple = (-1.0, -1.0, -1.0, True)
    func_indices: tuple = ()
    center_index: int = -1
    round1_indices: List[int] = field(default_factory=list)
    round2_indices: List[int] = field(default_factory=list)
    round3_indices: List[int] = field(default_factory=list)
    round4_indices: List[int] = field(default_factory=list)
    fully_linear: bool = False
    improving_directions: List[np.ndarray] = field(default_factory=list)
    # not part of the reference's RbfMeta: the round-4 factorisation kept on the device between prepare_update_model and
    # update_model (the reference recomputes it, RbfModel.jl:657-660), and the database size the selection saw
    _kept: object = field(default=None, repr=False, compare=False)


def _collect_indices(meta: RbfMeta, include_x: bool = True) -> List[int]:      # :178-186
    return (([meta.center_index] if include_x else []) + list(meta.round1_indices) + list(meta.round2_indices)
            + list(meta.round3_indices) + list(meta.round4_indices))


def get_saveable(meta: RbfMeta) -> RbfMeta:           # :166-175
    return RbfMeta(func_indices=None, center_index=meta.center_index, round1_indices=meta.round1_indices,
                   round2_indices=meta.round2_indices, round3_indices=meta.round3_indices,
                   round4_indices=meta.round4_indices, fully_linear=meta.fully_linear,
                   improving_directions=meta.improving_directions)


# ---------------------------------------------------------------------------------------- host objects the plugin reads
class ArrayDB:
    """The slice of ArrayDB the path touches (src/Databases.jl:15-32, 174-183, 202-212, 222-250, 258-277)."""

    def __init__(self, n_vars: int):
        self.n_vars = n_vars
        self._sites = np.zeros((64, n_vars))
        self.values: List[Optional[np.ndarray]] = []
        self.unevaluated_ids: List[int] = []
        self.num_entries = 0

    def sites_array(self) -> np.ndarray:
        return self._sites[: self.num_entries]

    def get_site(self, i: int) -> np.ndarray:
        return self._sites[i - 1]

    def get_value(self, i: int):
        return self.values[i - 1]

    def new_result(self, x, y=None) -> int:
        if self.num_entries == self._sites.shape[0]:
            self._sites = np.vstack((self._sites, np.zeros_like(self._sites)))
        new_id = self.num_entries + 1
        self._sites[self.num_entries] = np.asarray(x, dtype=np.float64)
        has_val = y is not None and len(y) > 0 and not np.any(np.isnan(y))
        self.values.append(np.array(y, dtype=np.float64) if has_val else None)
        if not has_val:
            self.unevaluated_ids.append(new_id)
        self.num_entries += 1
        return new_id

    def find_result(self, x) -> int:
        eq = np.all(self.sites_array() == np.asarray(x)[None, :], axis=1)
        hits = np.flatnonzero(eq)
        return int(hits[0]) + 1 if hits.size else -1

    def ensure_contains_res_with_site(self, x) -> int:
        pos = self.find_result(x)
        return pos if pos > 0 else self.new_result(x, None)

    def eval_missing(self, func) -> int:
        missing = list(self.unevaluated_ids)
        for i in missing:
            self.values[i - 1] = np.atleast_1d(np.asarray(func(self.get_site(i)), dtype=np.float64))
        for i in missing:
            self.unevaluated_ids.remove(i)
        return len(missing)


@dataclass
class SuperDB:
    """SuperDB, Databases.jl:340-401: one sub-database per group of function indices."""
    sub_dbs: Dict[tuple, ArrayDB]


def get_sub_db(sdb: SuperDB, func_indices) -> ArrayDB:
    return sdb.sub_dbs[tuple(func_indices)]


@dataclass
class IterData:
    """IterData fields the plugin reads (IterDataIterSaveable.jl:12-29)."""
    x_scaled: np.ndarray
    delta: float
    x_indices: Dict[tuple, int]


@dataclass
class VarScaler:
    """full_bounds_internal(scal): scaled global bounds, ±Inf when unbounded (VarScaler.jl:205-213)."""
    lb: np.ndarray
    ub: np.ndarray


@dataclass
class AlgoConfig:
    delta_max: float = float(np.float32(0.5))       # AbstractConfigInterface.jl:31 (Float32 literal)
    max_evals: int = INT_MAX
    # rtol of `Δ ≈ Δ_max` (RbfModel.jl:588).  Julia: max(sqrt(eps(T))) over the two argument types; the default config's
    # delta_max is Float32 => sqrt(eps(Float32)).  An AlgorithmConfig{Float64} gives sqrt(eps(Float64)) = 1.4901161193847656e-08.
    isapprox_rtol: float = float(np.sqrt(np.float32(np.finfo(np.float32).eps)))


@dataclass
class MopStub:
    """num_evals(_get(mop, ind)) per function index (RbfModel.jl:613)."""
    num_evals: Dict[object, int] = field(default_factory=dict)


# ---------------------------------------------------------------------------------------- model
class RbfModel:
    """RbfModel, RbfModel.jl:33-46: wraps the device-resident interpolation model."""

    def __init__(self, model: ModelBatch, fully_linear: bool = False):
        self.model = model
        self._fully_linear = bool(fully_linear)


def fully_linear(mod: RbfModel) -> bool:
    return mod._fully_linear


def set_fully_linear(mod: RbfModel, val: bool) -> None:
    mod._fully_linear = bool(val)


def num_outputs(mod: RbfModel) -> int:
    return mod.model.k


# ---------------------------------------------------------------------------------------- prepare_* (rounds 1-4)
def _exploit_other_rbf_metas(meta: RbfMeta, db: ArrayDB, sdb: SuperDB, meta_array) -> bool:
    """RbfModel.jl:311-342."""
    if meta_array is None:
        return False
    for other in meta_array:
        if isinstance(other, RbfMeta) and other.signature == meta.signature:
            other_db = get_sub_db(sdb, other.func_indices)
            for fn in ("round1_indices", "round2_indices", "round3_indices"):
                mine = getattr(meta, fn)
                mine.clear()
                for rid in getattr(other, fn):
                    mine.append(db.ensure_contains_res_with_site(other_db.get_site(rid)))
            meta.improving_directions = [d.copy() for d in other.improving_directions]
            meta.fully_linear = other.fully_linear
            return True
    return False


def _rbf_round4(db: ArrayDB, lb_2, ub_2, x, delta, indices_found_so_far: Sequence[int], cfg: RbfConfig,
                engine: Optional[Engine] = None) -> List[int]:
    """_rbf_round4, RbfModel.jl:352-499, as test/rbf_models.jl:74-86 calls it."""
    eng = engine or default_engine()
    cfg_num = _with_numeric_shape(cfg, delta)
    found = np.asarray(indices_found_so_far, dtype=np.int32)[None, :]
    r4, n_r4, status = eng.round4(cfg_num, db.sites_array()[None], [db.num_entries], np.asarray(lb_2)[None],
                                  np.asarray(ub_2)[None], found, [found.shape[1]])
    return [int(v) for v in r4[0, : n_r4[0]]]


def _with_numeric_shape(cfg: RbfConfig, delta: float) -> RbfConfig:
    if isinstance(cfg.shape_parameter, str):
        from dataclasses import replace
        return replace(cfg, shape_parameter=_shape_value(delta, cfg))
    return cfg


def prepare_init_model(cfg: RbfConfig, func_indices, mop, scal, iter_data, sdb, ac, *, ensure_fully_linear=True, **kw):
    """RbfModel.jl:506-513."""
    meta = RbfMeta(signature=cfg.signature(), func_indices=tuple(func_indices))
    return prepare_update_model(None, meta, cfg, func_indices, mop, scal, iter_data, sdb, ac,
                                ensure_fully_linear=ensure_fully_linear, **kw)


def prepare_update_model(mod, meta: RbfMeta, cfg: RbfConfig, func_indices, mop, scal: VarScaler, iter_data: IterData,
                         sdb: SuperDB, algo_config: AlgoConfig, *, ensure_fully_linear=False, force_rebuild=False,
                         meta_array=None, engine: Optional[Engine] = None) -> RbfMeta:
    """RbfModel.jl:518-655.  One C-ABI call does rounds 1-4; the database appends stay here."""
    eng = engine or default_engine()
    db = get_sub_db(sdb, func_indices)
    fit = tuple(func_indices)
    delta = float(iter_data.delta)
    delta_max = float(algo_config.delta_max)
    x = np.asarray(iter_data.x_scaled, dtype=np.float64)
    x_index = iter_data.x_indices[fit]
    n = len(x)
    meta.fully_linear = False
    skip_first_rounds = _exploit_other_rbf_metas(meta, db, sdb, meta_array)
    meta.center_index = x_index
    cfg_num = _with_numeric_shape(cfg, delta)
    if skip_first_rounds:                              # @goto round4, :562
        meta._kept = None
        meta.round4_indices = []
        if cfg.optimized_sampling:
            delta_2 = cfg.theta_enlarge_2 * delta_max
            lb_2, ub_2 = np.maximum(scal.lb, x - delta_2), np.minimum(scal.ub, x + delta_2)
            meta.round4_indices = _rbf_round4(db, lb_2, ub_2, x, delta, _collect_indices(meta), cfg, eng)
        return meta
    num_objf_evals = max((mop.num_evals.get(ind, 0) for ind in func_indices), default=0) if mop is not None else 0
    budget = min(algo_config.max_evals, cfg.max_evals) - 1 - num_objf_evals - len(db.unevaluated_ids)      # :613-618
    max_new = int(max(0, min(budget, 2**31 - 1)))
    n_db0 = db.num_entries
    prev = meta._kept[1] if meta._kept is not None else None
    eng.set_isapprox_rtol(getattr(algo_config, "isapprox_rtol", 1.4901161193847656e-08))
    res, prepared = eng.select_points_keep(cfg_num, db.sites_array()[None], [n_db0], [x_index], x[None], [delta], delta_max,
                                           scal.lb, scal.ub, ensure_fully_linear, force_rebuild, max_new, prepared=prev)
    meta._kept = (res, prepared, n_db0, x_index)
    meta.round1_indices = [int(v) for v in res.r1[0, : res.n_r1[0]]]
    meta.round2_indices = [int(v) for v in res.r2[0, : res.n_r2[0]]]
    meta.improving_directions = [res.dirs[0, c].copy() for c in range(res.n_dirs[0])]
    # round 3 sites become value-less results with consecutive ids (new_result!(db, p, F[]), :301-305)
    meta.round3_indices = [db.new_result(res.r3_sites[0, i], None) for i in range(res.n_r3[0])]
    meta.round4_indices = [int(v) for v in res.r4[0, : res.n_r4[0]]]
    meta.fully_linear = bool(res.flags_out[0, 0])
    return meta


def prepare_improve_model(mod, meta: RbfMeta, cfg: RbfConfig, func_indices, mop, scal: VarScaler, iter_data: IterData,
                          sdb: SuperDB, algo_config, **kw) -> RbfMeta:
    """RbfModel.jl:699-732 (one wall step along the first improving direction; n-sized host arithmetic)."""
    if not meta.fully_linear and meta.improving_directions:
        db = get_sub_db(sdb, func_indices)
        x = np.asarray(iter_data.x_scaled, dtype=np.float64)
        delta_1 = iter_data.delta * cfg.theta_enlarge_1
        lb_1, ub_1 = np.maximum(scal.lb, x - delta_1), np.minimum(scal.ub, x + delta_1)
        piv = delta_1 * cfg.theta_pivot
        d = meta.improving_directions.pop(0)
        length = _intersect_box_absmax(x, d, lb_1, ub_1)
        offset = length * d
        success = False
        if float(np.max(np.abs(offset))) > piv:
            meta.round1_indices.append(db.new_result(x + offset, None))
            meta._kept = None                       # the training set changed: the kept factorisation no longer describes it
            success = True
        if not meta.improving_directions and success:
            meta.fully_linear = True
    return meta


def _intersect_box_absmax(x, d, lb, ub) -> float:
    """intersect_box(...; return_vals=:absmax), utilities.jl:126-221 (host copy for prepare_improve_model)."""
    nz = d != 0
    if not np.any(nz):
        return math.inf
    sig = []
    for b, sense in ((lb, "lb"), (ub, "ub")):
        tmp = b[nz] - x[nz]
        dd = d[nz]
        z = tmp == 0
        with np.errstate(divide="ignore", invalid="ignore"):
            sig.append(tmp[~z] / dd[~z])
        if np.any(z):
            sig.append(np.where((dd[z] > 0) if sense == "lb" else (dd[z] < 0), np.inf, 0.0))
    sig = np.concatenate(sig)
    pos, neg = sig[sig >= 0], sig[~(sig >= 0)]
    s_pos = float(pos.min()) if pos.size else 0.0
    s_neg = float(neg.max()) if neg.size else 0.0
    return s_pos if abs(s_pos) >= abs(s_neg) else s_neg


# ---------------------------------------------------------------------------------------- build
def init_model(meta, cfg, func_indices, mop, scal, iter_data, sdb, ac, **kw):
    """RbfModel.jl:738-741."""
    return update_model(None, meta, cfg, func_indices, mop, scal, iter_data, sdb, ac, **kw)


def update_model(mod, meta: RbfMeta, cfg: RbfConfig, func_indices, mop, scal, iter_data: IterData, sdb: SuperDB, ac,
                 *, engine: Optional[Engine] = None, **kw) -> Tuple[RbfModel, RbfMeta]:
    """RbfModel.jl:743-767."""
    eng = engine or default_engine()
    db = get_sub_db(sdb, func_indices)
    ids = _collect_indices(meta)
    sites = np.array([db.get_site(i) for i in ids])
    vals = [db.get_value(i) for i in ids]
    if any(v is None for v in vals):
        raise ValueError("training results without values: call eval_missing! between prepare_* and update_model")
    values = np.array(vals)
    shape = _shape_value(iter_data.delta, cfg)
    cfg_num = _with_numeric_shape(cfg, iter_data.delta)
    if meta._kept is not None and not isinstance(cfg.shape_parameter, str):
        # round 4 kept its factorisation for exactly this training set: finish the model from it (two triangular solves)
        res, prepared, n_db0, x_index = meta._kept
        k = values.shape[1]
        dbv = np.zeros((1, n_db0, k))
        for i in range(n_db0):
            v = db.get_value(i + 1)
            if v is not None:
                dbv[0, i] = v
        r3v = np.zeros((1, sites.shape[1], k))                      # values of the new round-3 sites, n x k
        for j, rid in enumerate(meta.round3_indices):
            r3v[0, j] = db.get_value(rid)
        model, _status = eng.build_prepared(cfg_num, prepared, db.sites_array()[None, :n_db0], dbv, [x_index], res, r3v,
                                            recycle=mod.model if isinstance(mod, RbfModel) else None)
        return RbfModel(model, meta.fully_linear), meta
    model, _status = eng.build(cfg_num, sites[None], values[None], [len(ids)], [shape])
    return RbfModel(model, meta.fully_linear), meta


def improve_model(mod, meta, cfg, func_indices, mop, scal, iter_data, sdb, ac, **kw):
    """RbfModel.jl:770-776."""
    return update_model(mod, meta, cfg, func_indices, mop, scal, iter_data, sdb, ac, **kw)


# ---------------------------------------------------------------------------------------- evaluation
def eval_models(mod: RbfModel, scal, x_scaled, ell=None):
    """RbfModel.jl:783-790; `ell` may be an int or a list of ints (1-based), as RefSurrogate passes it."""
    Y, _ = mod.model.engine.eval(mod.model, np.asarray(x_scaled, dtype=np.float64)[None, None, :], True, False)
    y = Y[0, 0]
    if ell is None:
        return y
    return y[np.asarray(ell) - 1]


def get_jacobian(mod: RbfModel, scal, x_scaled, rows=None):
    """RbfModel.jl:798-800."""
    _, J = mod.model.engine.eval(mod.model, np.asarray(x_scaled, dtype=np.float64)[None, None, :], False, True)
    Jm = J[0, 0]
    return Jm if rows is None else Jm[np.asarray(rows) - 1, :]


def get_gradient(mod: RbfModel, scal, x_scaled, ell: int):
    """RbfModel.jl:793-795."""
    return get_jacobian(mod, scal, x_scaled, [ell])[0]


def _steepest_descent_direction(x, jac, lb, ub, A_eq=None, b_eq=None, A_ineq=None, b_ineq=None, normalize=True,
                                engine: Optional[Engine] = None):
    """_steepest_descent_direction, descent.jl:91-135: returns (d, omega).  The reference hands the LP to JuMP + OSQP
    (eps_rel = 1e-5); here it is solved exactly on the device.  Linear constraints of the MOP are outside the hot path."""
    for a in (A_eq, b_eq, A_ineq, b_ineq):
        if a is not None and len(a) > 0:
            raise NotImplementedError("linear constraints are not part of the GPU hot path")
    eng = engine or default_engine()
    d, omega, _it, status = eng.descent_direction(np.asarray(jac, dtype=np.float64)[None], np.asarray(x, dtype=np.float64)[None],
                                                  lb, ub, normalize)
    if status[0] != 0:           # the reference warns and returns (zeros, -Inf) when the LP solver fails (descent.jl:129-133)
        return np.zeros(len(x)), -math.inf
    return d[0], float(omega[0])


def _backtrack(x, direction, step_size, omega, mod: RbfModel, *, armijo_const_rhs=1e-6, armijo_const_shrink=0.75,
               min_stepsize=10 * np.finfo(np.float64).eps, max_loops=None, strict_backtracking=True):
    """_backtrack, descent.jl:150-185: returns (x₊, m(x₊), step)."""
    xp, mxp, step, _idx, _mx = mod.model.engine.backtrack(mod.model, np.asarray(x)[None], np.asarray(direction)[None],
                                                          [step_size], [omega], armijo_const_rhs, armijo_const_shrink,
                                                          min_stepsize, max_loops, strict_backtracking)
    return xp[0], mxp[0], step[0]


# ---------------------------------------------------------------------------------------- Pascoletti-Serafini
@dataclass
class PascolettiSerafiniConfig:
    """src/descent.jl:324-349 (only `:GN_ISRES`, the reference's default, runs on the device; a polish algorithm stays with NLopt)."""
    reference_point: Sequence[float] = ()
    reference_direction: Sequence[float] = ()
    trust_region_factor: float = 1.0
    max_ps_problem_evals: int = -1
    max_ps_polish_evals: int = -1
    max_ideal_point_problem_evals: int = -1
    main_algo: str = "GN_ISRES"
    reference_algo: str = "GN_ISRES"
    reference_trust_region_factor: float = 1.1
    ps_polish_algo: Optional[str] = None
    seed: int = 0

    def __post_init__(self):
        assert all(v > 0 for v in self.reference_direction), "The components of the `reference_direction` cannot be negative."
        if self.main_algo != "GN_ISRES" or self.reference_algo != "GN_ISRES":
            raise NotImplementedError("only :GN_ISRES inner solves run on the device")
        if self.ps_polish_algo is not None:
            raise NotImplementedError("ps_polish_algo is an NLopt local method of the reference's host code")


def _get_global_dir(cfg: PascolettiSerafiniConfig, fx):
    """descent.jl:359-367."""
    if len(cfg.reference_direction) > 0:
        return np.asarray(cfg.reference_direction, dtype=np.float64)
    if len(cfg.reference_point) > 0:
        return np.asarray(fx, dtype=np.float64) - np.asarray(cfg.reference_point, dtype=np.float64)
    return None


def _local_bounds(x, delta, lb, ub):
    """local_bounds(scal, x, delta), utilities.jl:290-294."""
    x = np.asarray(x, dtype=np.float64)
    return np.maximum(np.broadcast_to(lb, x.shape), x - delta), np.minimum(np.broadcast_to(ub, x.shape), x + delta)


def compute_local_ideal_point(mod: RbfModel, x_scaled, lb_eff, ub_eff, max_evals: int = -1, n_obj: Optional[int] = None, seed: int = 0):
    """compute_local_ideal_point, descent.jl:404-412: one box-constrained minimisation per objective (`_min_component` :369-387)."""
    eng, k = mod.model.engine, mod.model.k
    n_obj = k if n_obj is None else n_obj
    out = np.zeros(n_obj)
    for ell in range(n_obj):
        f, _x, _y, found, _ = eng.ps_solve(mod.model, np.asarray(x_scaled)[None], np.asarray(lb_eff)[None], np.asarray(ub_eff)[None],
                                           None, None, n_obj, ell, -1, max_evals, seed + 7919 * (ell + 1))
        out[ell] = f[0] if found[0] else math.inf
    return out


def get_criticality_ps(desc_cfg: PascolettiSerafiniConfig, mod: RbfModel, x_scaled, fx, delta, lb, ub, n_obj: Optional[int] = None):
    """get_criticality(::PascolettiSerafiniConfig, ...), descent.jl:503-581, for one RBF surrogate group holding the objectives
    (and, behind them, constraint surrogates c <= 0).  Returns (omega, x_trial, m(x_trial), step length) -- the reference's
    `(0, copy(x), mx, 0)` when the direction has a non-positive component (:541-543) or the solver reports no feasible point."""
    x = np.asarray(x_scaled, dtype=np.float64)
    eng, n, k = mod.model.engine, len(x), mod.model.k
    n_obj = k if n_obj is None else n_obj
    lb_eff, ub_eff = _local_bounds(x, delta, lb, ub)
    mx = eval_models(mod, None, x)
    r = _get_global_dir(desc_cfg, fx)
    if r is None:
        ideal = compute_local_ideal_point(mod, x, lb_eff, ub_eff, desc_cfg.max_ideal_point_problem_evals, n_obj, desc_cfg.seed)
        r = np.asarray(fx, dtype=np.float64)[:n_obj] - ideal
    if np.any(r[:n_obj] <= 0):
        return 0.0, x.copy(), mx, 0.0
    rr = np.ones(k); rr[:n_obj] = r[:n_obj]
    max_evals = 500 * (n + 1) if desc_cfg.max_ps_problem_evals < 0 else desc_cfg.max_ps_problem_evals       # _ps_max_evals, :414-433
    tau, xm, ym, found, _ = eng.ps_solve(mod.model, x[None], lb_eff[None], ub_eff[None], mx[None], rr[None], n_obj, -1, -1, max_evals,
                                         desc_cfg.seed)
    if not found[0] or not np.isfinite(tau[0]) or np.any(np.isnan(xm[0])):
        return 0.0, x.copy(), mx, 0.0
    return float(abs(tau[0])), xm[0], ym[0], float(np.max(np.abs(x - xm[0])))
