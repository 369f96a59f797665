"""CPU oracle for Morbit.jl's RBF-surrogate hot path (NumPy + SciPy/LAPACK, float64).

TEST INFRASTRUCTURE ONLY.  Nothing under ``morbit.jl_b200/`` imports this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may use it,
and only as the checker.

PARITY UNPINNED (by the reference): the reference is pure Julia, no ``julia`` binary exists
in this image or on the GPU boxes, and the reference ships no golden vectors for this path
(test/rbf_models.jl holds properties only).  This file is therefore a literal
restatement of the reference sources, pinned by the properties of test/rbf_models.jl
(tests/test_oracle_properties.py), by cross-checking against the independent C restatement
in ``oracle/rbf_oracle.c`` (tests/test_oracle_c_vs_py.py) and -- for the model build and
evaluation, whose arithmetic lives in the un-vendored dependency -- by an implementation
written by other people: ``scipy.interpolate.RBFInterpolator`` (tests/test_scipy_pin.py:
seven radial functions, tails of degree -1/0/1, values to cond * eps, Jacobians against
finite differences of SciPy's values).

Half of the arithmetic lives in the un-vendored dependency RadialBasisFunctionModels.jl
(compat "0.3.4", Project.toml:25,50; no Manifest, so no exact pin).  Its published
algorithm is restated here from the call sites in src/models/RbfModel.jl
(:374, :487, :695, :759-763, :784-799); every convention that mathematics does not fix
is listed in DESIGN.md ("assumption register", U1-U9).

All ids are 1-based like the reference's Int ids; arrays hold them as Python ints.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import scipy.linalg as sla

EPS = float(np.finfo(np.float64).eps)
INT_MAX = 2**63 - 1

# src/models/RbfModel.jl:48-54
RBF_KERNELS = ("cubic", "inv_multiquadric", "multiquadric", "thin_plate_spline", "gaussian")


# --------------------------------------------------------------------------------------
# A1  RbfConfig                                                  src/models/RbfModel.jl:66-112
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class RbfConfig:
    kernel: str = "cubic"
    shape_parameter: float = float("nan")      # String shape parameters are host-side sugar
    polynomial_degree: int = 1
    theta_enlarge_1: float = 2.0
    theta_enlarge_2: float = 2.0
    theta_pivot: Optional[float] = None         # default 1/(2 theta_enlarge_1), :83
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
        # asserts :102-111
        assert self.theta_enlarge_1 * self.theta_pivot <= 1
        assert self.kernel in RBF_KERNELS
        if self.kernel == "thin_plate_spline":
            assert math.isnan(sp) or (sp % 1 == 0 and sp >= 1)
        if self.kernel == "cubic":
            assert math.isnan(sp) or (sp % 1 == 0 and sp % 2 == 1)
        assert math.isnan(sp) or sp > 0
        assert self.theta_enlarge_1 >= 1 and self.theta_enlarge_2 >= 1

    def signature(self):                         # _get_signature, :114
        return (self.theta_pivot, self.theta_enlarge_1, self.theta_enlarge_2, self.optimized_sampling)


# --------------------------------------------------------------------------------------
# A13  radial functions  (RBF._get_rad_func, called at RbfModel.jl:695; dependency restated)
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class RadialFunction:
    """phi(rho) and psi(rho) = phi'(rho)/rho for one kernel with resolved parameters."""
    kernel: str
    alpha: float      # shape parameter (gaussian, (inv_)multiquadric)
    beta: float       # exponent (cubic: odd int; multiquadrics: 1/2; tps: k)

    @property
    def cpd_order(self) -> int:
        if self.kernel == "cubic":
            return int(math.ceil(self.beta / 2))
        if self.kernel == "multiquadric":
            return int(math.ceil(self.beta))
        if self.kernel == "thin_plate_spline":
            return int(self.beta) + 1
        return 0

    def phi(self, rho):
        rho = np.asarray(rho, dtype=np.float64)
        k, a, b = self.kernel, self.alpha, self.beta
        if k == "cubic":
            sgn = (-1.0) ** math.ceil(b / 2)
            return sgn * rho ** b
        if k == "multiquadric":
            sgn = (-1.0) ** math.ceil(b)
            return sgn * (1.0 + (a * rho) ** 2) ** b
        if k == "inv_multiquadric":
            return (1.0 + (a * rho) ** 2) ** (-b)
        if k == "gaussian":
            return np.exp(-((a * rho) ** 2))
        if k == "thin_plate_spline":
            kk = int(b)
            sgn = (-1.0) ** (kk + 1)
            with np.errstate(divide="ignore", invalid="ignore"):
                out = sgn * rho ** (2 * kk) * np.log(rho)
            return np.where(rho == 0, 0.0, out)
        raise ValueError(k)

    def psi(self, rho):
        """phi'(rho)/rho, with the rho == 0 contribution defined as 0 where singular (U6)."""
        rho = np.asarray(rho, dtype=np.float64)
        k, a, b = self.kernel, self.alpha, self.beta
        if k == "cubic":
            sgn = (-1.0) ** math.ceil(b / 2)
            with np.errstate(divide="ignore", invalid="ignore"):
                out = sgn * b * rho ** (b - 2)
            return np.where(rho == 0, 0.0, out)
        if k == "multiquadric":
            sgn = (-1.0) ** math.ceil(b)
            return sgn * 2 * b * a * a * (1.0 + (a * rho) ** 2) ** (b - 1)
        if k == "inv_multiquadric":
            return -2 * b * a * a * (1.0 + (a * rho) ** 2) ** (-b - 1)
        if k == "gaussian":
            return -2 * a * a * np.exp(-((a * rho) ** 2))
        if k == "thin_plate_spline":
            kk = int(b)
            sgn = (-1.0) ** (kk + 1)
            with np.errstate(divide="ignore", invalid="ignore"):
                out = sgn * rho ** (2 * kk - 2) * (2 * kk * np.log(rho) + 1.0)
            return np.where(rho == 0, 0.0, out)
        raise ValueError(k)


def get_radial_function(cfg: RbfConfig, shape: Optional[float] = None) -> RadialFunction:
    """_get_kernel_params + _get_rad_func, RbfModel.jl:665-696.  NaN -> package defaults (:673)."""
    sp = cfg.shape_parameter if shape is None else shape
    nan = math.isnan(sp)
    k = cfg.kernel
    if k == "gaussian":
        return RadialFunction(k, 1.0 if nan else sp, 0.0)
    if k in ("multiquadric", "inv_multiquadric"):
        return RadialFunction(k, 1.0 if nan else sp, 0.5)          # (sp, 1//2), :680-682
    if k == "cubic":
        return RadialFunction(k, 1.0, 3.0 if nan else float(int(sp)))   # Int(sp), :684
    if k == "thin_plate_spline":
        return RadialFunction(k, 1.0, 2.0 if nan else float(int(sp)))   # Int(sp), :686
    raise ValueError(k)


def poly_basis(x: np.ndarray, degree: int) -> np.ndarray:
    """Canonical monomial basis of total degree <= degree: [], [1], [1, x_1..x_n] or [1, x_1..x_n, x_i x_j (i <= j, i outer)]  (U4).
    Degree 2 only arises when a kernel's order of conditional positive definiteness raises it (thin plate spline k = 2, cubic beta = 5):
    the configuration itself allows -1..1 (RbfModel.jl:78)."""
    if degree < 0:
        return np.zeros(0)
    if degree == 0:
        return np.ones(1)
    if degree == 1:
        return np.concatenate(([1.0], x))
    n = len(x)
    quad = [x[i] * x[j] for i in range(n) for j in range(i, n)]
    return np.concatenate(([1.0], x, quad))


def poly_jac(x: np.ndarray, degree: int) -> np.ndarray:
    """d poly_basis / d x : (dim, n)."""
    n = len(x)
    if degree < 0:
        return np.zeros((0, n))
    if degree == 0:
        return np.zeros((1, n))
    rows = [np.zeros(n)] + [np.eye(n)[i] for i in range(n)]
    if degree >= 2:
        for i in range(n):
            for j in range(i, n):
                g = np.zeros(n); g[i] += x[j]; g[j] += x[i]
                rows.append(g)
    return np.array(rows)


def poly_dim(n: int, degree: int) -> int:
    return 0 if degree < 0 else (1 if degree == 0 else (n + 1 if degree == 1 else ((n + 1) * (n + 2)) // 2))


# --------------------------------------------------------------------------------------
# Database bookkeeping                       src/Databases.jl:15-32, 174-183, 202-212, 222-250
# --------------------------------------------------------------------------------------
class ArrayDB:
    def __init__(self):
        self.sites: List[np.ndarray] = []
        self.values: List[Optional[np.ndarray]] = []
        self.unevaluated_ids: List[int] = []

    @property
    def num_entries(self) -> int:
        return len(self.sites)

    def get_site(self, i: int) -> np.ndarray:
        return self.sites[i - 1]

    def get_value(self, i: int):
        return self.values[i - 1]

    def new_result(self, x, y=None) -> int:              # new_result!, :174-183
        new_id = self.num_entries + 1
        self.sites.append(np.array(x, dtype=np.float64))
        has_val = y is not None and len(y) > 0 and not np.any(np.isnan(y))
        self.values.append(np.array(y, dtype=np.float64) if has_val else None)
        if not has_val:
            self.unevaluated_ids.append(new_id)          # set_evaluated_flag!(…, false), :202-205
        return new_id

    def find_result(self, x) -> int:                     # :222-230 (site only)
        for i, s in enumerate(self.sites):
            if np.array_equal(s, x):
                return i + 1
        return -1

    def ensure_contains_res_with_site(self, x) -> int:   # :243-250
        pos = self.find_result(x)
        if pos < 0:
            pos = self.new_result(x, None)
        return pos

    def eval_missing(self, func):                        # eval_missing!, :258-277
        missing = list(self.unevaluated_ids)
        for i in missing:
            self.values[i - 1] = np.atleast_1d(np.asarray(func(self.sites[i - 1]), dtype=np.float64))
        for i in missing:
            self.unevaluated_ids.remove(i)
        return len(missing)


def results_in_box_indices(db: ArrayDB, lb, ub, exclude: Sequence[int] = ()) -> List[int]:
    """A3, src/Databases.jl:324-327: ascending ids, inclusive bounds, unevaluated sites count."""
    ex = set(exclude)
    out = []
    for i in range(1, db.num_entries + 1):
        s = db.sites[i - 1]
        if i not in ex and bool(np.all((lb <= s) & (s <= ub))):
            out.append(i)
    return out


# --------------------------------------------------------------------------------------
# Geometry helpers                                            src/utilities.jl:126-221, 285-300
# --------------------------------------------------------------------------------------
def local_bounds(x, delta, glb, gub):
    """_local_bounds, utilities.jl:290-294."""
    return np.maximum(glb, x - delta), np.minimum(gub, x + delta)


def _intersect_bound_vec(x, b, d, d_nz, sense):
    """utilities.jl:126-152."""
    dd = d[d_nz]
    tmp = b[d_nz] - x[d_nz]
    tmp_z = tmp == 0
    with np.errstate(divide="ignore", invalid="ignore"):
        sig = tmp[~tmp_z] / dd[~tmp_z]
    _d = dd[tmp_z]
    if _d.size == 0:
        return sig
    if sense == "lb":
        on = np.where(_d > 0, np.inf, 0.0)
    else:
        on = np.where(_d < 0, np.inf, 0.0)
    return np.concatenate((sig, on))


WALL_TIES: list = []      # (s_pos, s_neg) of every intersect_box call whose two wall distances agreed to 1e-9 (knife-edge sign); tests clear it


def intersect_box_absmax(x, d, lb, ub) -> float:
    """intersect_box(...; return_vals=:absmax) -> _intersect_bounds, utilities.jl:156-221, 285-287."""
    if not np.any(d != 0):
        return math.inf
    d_nz = d != 0
    sig = np.concatenate((_intersect_bound_vec(x, lb, d, d_nz, "lb"),
                          _intersect_bound_vec(x, ub, d, d_nz, "ub")))
    if sig.size == 0:
        return math.inf
    pos = sig[sig >= 0]
    neg = sig[~(sig >= 0)]
    s_pos = float(pos.min()) if pos.size else 0.0
    s_neg = float(neg.max()) if neg.size else 0.0
    if pos.size and neg.size and abs(abs(s_pos) - abs(s_neg)) <= 1e-9 * max(abs(s_pos), abs(s_neg)):
        WALL_TIES.append((s_pos, s_neg))                      # the sign of the step is decided by the last bits of x, d (test hook)
    return s_pos if abs(s_pos) >= abs(s_neg) else s_neg       # positive wins ties, :212-217


def givens(f: float, g: float) -> Tuple[float, float, float]:
    """LinearAlgebra.givensAlgorithm (LAPACK dlartg, pre-3.10 convention); SURVEY App. A.4."""
    if g == 0:
        return 1.0, 0.0, f
    if f == 0:
        return 0.0, 1.0, g
    r = math.hypot(f, g)
    c, s = f / r, g / r
    if abs(f) > abs(g) and c < 0:
        c, s, r = -c, -s, -r
    return c, s, r


def nullify_last_row(R: np.ndarray):
    """utilities.jl:437-448: dense G with G @ R_in == R_out, last row annihilated by Givens."""
    R = R.copy()
    m, n = R.shape
    G = np.eye(m)
    for j in range(min(m - 1, n)):
        c, s, _ = givens(R[j, j], R[m - 1, j])
        g = np.eye(m)
        g[j, j] = c; g[j, m - 1] = s; g[m - 1, j] = -s; g[m - 1, m - 1] = c
        R = g @ R
        G = g @ G
    return R, G


# --------------------------------------------------------------------------------------
# A4  affinely independent point filter            src/models/AffinelyIndependentPoints.jl
# --------------------------------------------------------------------------------------
def orthogonal_complement_matrix(Y: np.ndarray) -> np.ndarray:
    """:4-11  full Householder QR (LAPACK geqrf/orgqr), trailing columns, inf-norm scaled."""
    n, j = Y.shape
    Q, _ = sla.qr(Y, mode="full")
    Z = Q[:, j:].copy()
    if Z.shape[1] > 0:
        Z /= np.max(np.abs(Z), axis=0)[None, :]
    return Z


@dataclass
class FilterTrace:
    """Decision margins, so tests can detect knife-edge inputs."""
    best: List[float] = field(default_factory=list)
    runner_up: List[float] = field(default_factory=list)
    first: List[bool] = field(default_factory=list)     # entry belongs to an unconditional first pick (exact arg-max of exact norms)

    def knife_edge(self, rtol: float = 1e-9) -> bool:
        """True when a pivot test or an arg-max between candidates was decided by less than rtol (relative)."""
        return any((not f) and abs(b - r) <= rtol * max(abs(b), abs(r)) for b, r, f in zip(self.best, self.runner_up, self.first)
                   if math.isfinite(b) and math.isfinite(r))


def affinely_independent_filter(x0, seeds: Sequence[np.ndarray], pivot_val: float, n_wanted: int,
                                Y: Optional[np.ndarray] = None, Z: Optional[np.ndarray] = None,
                                trace: Optional[FilterTrace] = None):
    """:51-106.  Returns (positions into seeds (0-based), Y, Z)."""
    n = len(x0)
    Y = np.zeros((n, 0)) if Y is None else Y.copy()
    Z = np.eye(n) if Z is None else Z.copy()
    shifted = [s - x0 for s in seeds]
    picked: List[int] = []
    if len(shifted) == 0:
        return picked, Y, Z
    # first iterate (:51-69): argmax inf-norm, first maximiser, accepted unconditionally
    norms = [float(np.max(np.abs(s))) for s in shifted]
    i0 = int(np.argmax(norms))                   # np.argmax returns the first maximiser
    if trace is not None:
        srt = sorted(norms, reverse=True)
        trace.best.append(srt[0]); trace.runner_up.append(srt[1] if len(srt) > 1 else -math.inf); trace.first.append(True)
    Y = np.hstack((Y, shifted[i0][:, None]))
    Z = orthogonal_complement_matrix(Y)
    cand = [i for i in range(len(shifted)) if i != i0]
    picked.append(i0)
    num_found = 1
    while True:                                   # next iterates (:71-106)
        if num_found == n_wanted or len(cand) == 0:
            break
        best_val, best_index, second = -math.inf, -1, -math.inf
        for i in cand:
            val = float(np.max(np.abs(Z @ (Z.T @ shifted[i])))) if Z.shape[1] > 0 else 0.0
            if val > best_val:
                second = best_val
                best_val, best_index = val, i
            elif val > second:
                second = val
        if trace is not None:
            trace.best.append(best_val); trace.runner_up.append(max(second, pivot_val)); trace.first.append(False)
        if best_val > pivot_val:
            Y = np.hstack((Y, shifted[best_index][:, None]))
            Z = orthogonal_complement_matrix(Y)
            cand.remove(best_index)
            picked.append(best_index)
            num_found += 1
        else:
            break
    return picked, Y, Z


# --------------------------------------------------------------------------------------
# A2  RbfMeta                                                   src/models/RbfModel.jl:148-186
# --------------------------------------------------------------------------------------
@dataclass
class RbfMeta:
    signature: tuple = (-1.0, -1.0, -1.0, True)
    center_index: int = -1
    round1_indices: List[int] = field(default_factory=list)
    round2_indices: List[int] = field(default_factory=list)
    round3_indices: List[int] = field(default_factory=list)
    round4_indices: List[int] = field(default_factory=list)
    fully_linear: bool = False
    improving_directions: List[np.ndarray] = field(default_factory=list)

    def collect_indices(self) -> List[int]:          # _collect_indices, :178-186
        return ([self.center_index] + list(self.round1_indices) + list(self.round2_indices)
                + list(self.round3_indices) + list(self.round4_indices))


def _find_suitable_points(db, lb, ub, x, x_index, piv, already=(), Y=None, Z=None, n_missing=None,
                          trace=None):
    """RbfModel.jl:205-238."""
    n = len(x)
    cand = results_in_box_indices(db, lb, ub, [x_index] + list(already))
    pos, Y2, Z2 = affinely_independent_filter(x, [db.get_site(i) for i in cand], piv,
                                              n if n_missing is None else n_missing, Y, Z, trace)
    filtered = [cand[p] for p in pos]
    dirs = [Z2[:, c].copy() for c in range(Z2.shape[1])][::-1]      # reverse(eachcol(Z)), :232
    return filtered, dirs, cand, Y2, Z2


def _rbf_round3(db, lb_1, ub_1, x, piv, dirs, max_new, n_missing, ensure_fully_linear, force_rebuild):
    """RbfModel.jl:269-307."""
    n_new = max(0, min(n_missing, max_new))
    fully_linear = n_new >= n_missing
    assert len(dirs) >= n_new
    new_points = []
    for i in range(n_new):
        d = dirs[i]
        length = intersect_box_absmax(x, d, lb_1, ub_1)
        offset = length * d
        if float(np.max(np.abs(offset))) <= piv:
            if ensure_fully_linear and not force_rebuild:
                return None, None, None
            fully_linear = False
        new_points.append(x + offset)
    new_ids = [db.new_result(p, None) for p in new_points]
    return new_ids, fully_linear, dirs[n_new:]


@dataclass
class Round4Trace:
    tau2: List[float] = field(default_factory=list)
    accepted: List[bool] = field(default_factory=list)
    cid: List[int] = field(default_factory=list)          # candidate id each tau2 belongs to
    n_points: List[int] = field(default_factory=list)     # model points N when the candidate was tested
    sigma: List[float] = field(default_factory=list)      # the two terms tau2 = sigma - lv2 is the difference of
    lv2: List[float] = field(default_factory=list)
    guard: List[bool] = field(default_factory=list)       # candidate skipped by the rank guard (:433-438); tau2 holds the row norm
    phi_scale: float = 1.0                                # max |Phi_ij| of the starting kernel matrix (noise scale of tau2)


def rbf_round4(db: ArrayDB, lb_2, ub_2, x, delta, found: Sequence[int], cfg: RbfConfig,
               trace: Optional[Round4Trace] = None, return_state: bool = False):
    """A8, RbfModel.jl:352-499 -- literal dense restatement (O(N^3) per accepted point).

    ``use_max_points`` draws random points (:409-414) and is excluded from parity.
    """
    n = len(x)
    max_points = ((n + 1) * (n + 2)) // 2 if cfg.max_model_points <= 0 else cfg.max_model_points
    N = len(found)
    cand = results_in_box_indices(db, lb_2, ub_2, found)
    r4: List[int] = []
    state = None
    if N < max_points and (len(cand) > 0 or cfg.use_max_points):
        if cfg.use_max_points:
            raise NotImplementedError("use_max_points draws rand points; outside the parity corpus")
        chol_pivot = cfg.theta_pivot_cholesky ** 2                       # :370
        centers = [db.get_site(i).copy() for i in found]
        rf = get_radial_function(cfg)
        deg = cfg.polynomial_degree
        C = np.array(centers)
        Phi = rf.phi(np.sqrt(((C[:, None, :] - C[None, :, :]) ** 2).sum(-1)))
        Pi = np.array([poly_basis(c, deg) for c in centers]).reshape(N, poly_dim(n, deg))
        p = Pi.shape[1]
        if p > 0:
            Q, R = sla.qr(Pi, mode="full")                               # :381-389
        else:
            Q, R = np.eye(N), np.zeros((N, 0))
        Z = Q[:, N:]                                                     # :391 (always N x 0)
        L = np.zeros((0, 0)); Linv = np.zeros((0, 0))                    # :394-396
        phi0 = Phi[0, 0]                                                 # :398
        full_rank_dim = math.comb(n + deg, n) if deg >= 0 else 0         # binomial(n+deg, n), :433
        while N < max_points and cand:                                   # :402
            cid = cand.pop(0)                                            # :405
            xi = db.get_site(cid)
            phi_xi = rf.phi(np.sqrt(((C - xi[None, :]) ** 2).sum(-1)))   # kernels(xi), :421
            pi_xi = poly_basis(xi, deg)                                  # :424
            R_xi = np.vstack((R, pi_xi[None, :]))
            R_xi, G = nullify_last_row(R_xi)                             # :431
            if N < full_rank_dim:
                if np.linalg.norm(R_xi[-1, :]) <= EPS * 10:             # :434
                    if trace is not None:
                        trace.tau2.append(float(np.linalg.norm(R_xi[-1, :]))); trace.accepted.append(False); trace.cid.append(int(cid))
                        trace.n_points.append(int(N)); trace.sigma.append(0.0); trace.lv2.append(0.0); trace.guard.append(True)
                    continue
            g_t = G.T[:-1, -1]                                           # :442
            g_h = G[-1, -1]                                              # :443
            Qg = Q @ g_t
            v = Z.T @ (Phi @ Qg + phi_xi * g_h)                          # :446
            sigma = Qg @ Phi @ Qg + (2 * g_h) * (phi_xi @ Qg) + g_h ** 2 * phi0   # :447
            lv2 = float(np.linalg.norm(Linv @ v)) ** 2
            tau2 = sigma - lv2                                           # :449
            ok = tau2 > chol_pivot ** 2                                  # :452 (squared twice)
            if trace is not None:
                trace.tau2.append(float(tau2)); trace.accepted.append(bool(ok))
                trace.cid.append(int(cid)); trace.n_points.append(int(N))
                trace.sigma.append(float(sigma)); trace.lv2.append(lv2); trace.guard.append(False)
                trace.phi_scale = max(trace.phi_scale, float(np.max(np.abs(Phi))))
            if ok:
                r4.append(cid)
                tau = math.sqrt(tau2)
                Qa = np.zeros((N + 1, N + 1)); Qa[:N, :N] = Q; Qa[N, N] = 1.0
                Q = Qa @ G.T                                             # :462
                m = Z.shape[1]
                Z = np.block([[Z, Qg[:, None]], [np.zeros((1, m)), np.array([[g_h]])]])   # :464-467
                L = np.block([[L, np.zeros((m, 1))], [(v @ Linv.T)[None, :], np.array([[tau]])]])
                Linv = np.block([[Linv, np.zeros((m, 1))],
                                 [(-(v @ Linv.T @ Linv) / tau)[None, :], np.array([[1 / tau]])]])
                R = R_xi                                                 # :479
                Phi = np.block([[Phi, phi_xi[:, None]], [phi_xi[None, :], np.array([[phi0]])]])
                C = np.vstack((C, xi[None, :]))
                N += 1
        state = dict(Q=Q, R=R, Z=Z, L=L, Linv=Linv, Phi=Phi)
    if return_state:
        return r4, state
    return r4


ISAPPROX_RTOL_F64 = math.sqrt(EPS)
ISAPPROX_RTOL_F32 = math.sqrt(float(np.finfo(np.float32).eps))      # 3.4526698e-4


def isapprox(a: float, b: float, rtol: float = ISAPPROX_RTOL_F64) -> bool:
    """Julia's default isapprox: rtol = max(sqrt(eps(T)) over the argument types).  At RbfModel.jl:588 the radius is compared with
    delta_max(algo_config), a Float32 literal in the default config (AbstractConfigInterface.jl:31) => sqrt(eps(Float32));
    a user-supplied AlgorithmConfig{Float64} gives sqrt(eps(Float64))."""
    return abs(a - b) <= rtol * max(abs(a), abs(b))


def prepare_update_model(meta: RbfMeta, cfg: RbfConfig, db: ArrayDB, x, x_index: int, delta: float,
                         delta_max: float, glb, gub, *, ensure_fully_linear=False, force_rebuild=False,
                         meta_array: Optional[Sequence[Tuple[RbfMeta, ArrayDB]]] = None,
                         num_objf_evals: int = 0, algo_max_evals: int = INT_MAX,
                         trace: Optional[FilterTrace] = None, trace4: Optional[Round4Trace] = None,
                         isapprox_rtol: float = ISAPPROX_RTOL_F64):
    """A9, RbfModel.jl:518-655 (state machine of SURVEY App. A.2)."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    meta.fully_linear = False                                            # :541
    skip = False
    if meta_array is not None:                                           # _exploit_other_rbf_metas!, :311-342
        for other, other_db in meta_array:
            if other.signature == meta.signature:
                for fn in ("round1_indices", "round2_indices", "round3_indices"):
                    mine = getattr(meta, fn)
                    mine.clear()
                    for rid in getattr(other, fn):
                        mine.append(db.ensure_contains_res_with_site(other_db.get_site(rid)))
                meta.improving_directions = [d.copy() for d in other.improving_directions]
                meta.fully_linear = other.fully_linear
                skip = True
                break
    meta.center_index = x_index                                          # :548
    delta_1 = cfg.theta_enlarge_1 * delta
    lb_1, ub_1 = local_bounds(x, delta_1, glb, gub)
    piv = cfg.theta_pivot * delta_1
    delta_2 = cfg.theta_enlarge_2 * delta_max
    lb_2, ub_2 = local_bounds(x, delta_2, glb, gub)

    if not skip:
        if force_rebuild or not cfg.optimized_sampling:                  # :564-569
            filtered_1, cand_1 = [], []
            dirs = [np.eye(n)[:, i].copy() for i in range(n)]
            Y_1 = Z_1 = None
        else:
            filtered_1, dirs, cand_1, Y_1, Z_1 = _find_suitable_points(db, lb_1, ub_1, x, x_index, piv,
                                                                       trace=trace)
        meta.round1_indices = list(filtered_1)
        meta.improving_directions = list(dirs)
        n_missing = n - len(meta.round1_indices)
        if (n_missing == 0 or force_rebuild or not cfg.optimized_sampling or ensure_fully_linear
                or (isapprox(delta, delta_max, isapprox_rtol) and cfg.theta_enlarge_1 == cfg.theta_enlarge_2)):   # :588
            meta.fully_linear = True
            meta.round2_indices = []
        else:
            filtered_2, _, _, _, _ = _find_suitable_points(db, lb_2, ub_2, x, x_index, piv, already=cand_1,
                                                           Y=Y_1, Z=Z_1, n_missing=n_missing, trace=trace)
            meta.round2_indices = list(filtered_2)
        n_missing -= len(meta.round2_indices)
        meta.round3_indices = []
        if n_missing > 0:
            max_new = max(0, min(algo_max_evals, cfg.max_evals) - 1 - num_objf_evals
                          - len(db.unevaluated_ids))                     # :613-618
            new_ids, fl, _rest = _rbf_round3(db, lb_1, ub_1, x, piv, dirs, max_new, n_missing,
                                             ensure_fully_linear, force_rebuild)
            if new_ids is not None:
                meta.round3_indices = list(new_ids)
                meta.fully_linear = bool(fl) and len(meta.round2_indices) == 0       # :631
            else:
                return prepare_update_model(meta, cfg, db, x, x_index, delta, delta_max, glb, gub,
                                            ensure_fully_linear=True, force_rebuild=True,
                                            num_objf_evals=num_objf_evals, algo_max_evals=algo_max_evals,
                                            trace=trace, trace4=trace4, isapprox_rtol=isapprox_rtol)  # :634-637
    meta.round4_indices = []
    if cfg.optimized_sampling:                                           # :647-652
        meta.round4_indices = rbf_round4(db, lb_2, ub_2, x, delta, meta.collect_indices(), cfg, trace4)
    return meta


def prepare_improve_model(meta: RbfMeta, cfg: RbfConfig, db: ArrayDB, x, delta: float, glb, gub):
    """RbfModel.jl:699-732."""
    if not meta.fully_linear and meta.improving_directions:
        x = np.asarray(x, dtype=np.float64)
        delta_1 = delta * cfg.theta_enlarge_1
        lb_1, ub_1 = local_bounds(x, delta_1, glb, gub)
        piv = delta_1 * cfg.theta_pivot
        d = meta.improving_directions.pop(0)
        length = intersect_box_absmax(x, d, lb_1, ub_1)
        offset = length * d
        success = False
        if float(np.max(np.abs(offset))) > piv:
            meta.round1_indices.append(db.new_result(x + offset, None))
            success = True
        if not meta.improving_directions and success:
            meta.fully_linear = True
    return meta


# --------------------------------------------------------------------------------------
# A10/A11  interpolation model          RBF.RBFInterpolationModel etc. (dependency restated)
# --------------------------------------------------------------------------------------
@dataclass
class RbfModel:
    centers: np.ndarray        # N x n
    w: np.ndarray              # N x k
    lam: np.ndarray            # p x k
    rf: RadialFunction
    degree: int
    fully_linear: bool = False
    cond: float = float("nan")

    @property
    def num_outputs(self) -> int:
        return self.w.shape[1]

    def eval(self, x, ell=None):
        """eval_models, RbfModel.jl:783-790."""
        x = np.asarray(x, dtype=np.float64)
        rho = np.sqrt(((self.centers - x[None, :]) ** 2).sum(-1))
        y = self.rf.phi(rho) @ self.w
        if self.lam.shape[0] > 0:
            y = y + poly_basis(x, self.degree) @ self.lam
        return y if ell is None else y[np.asarray(ell) - 1]

    def jac(self, x, rows=None):
        """get_jacobian, RbfModel.jl:798-800:  rows x n."""
        x = np.asarray(x, dtype=np.float64)
        diff = x[None, :] - self.centers
        rho = np.sqrt((diff ** 2).sum(-1))
        J = (self.w * self.rf.psi(rho)[:, None]).T @ diff                # k x n
        if self.degree == 1:
            J = J + self.lam[1:, :].T
        elif self.degree >= 2:
            J = J + self.lam.T @ poly_jac(x, self.degree)
        return J if rows is None else J[np.asarray(rows) - 1, :]

    def grad(self, x, ell: int):
        """get_gradient, RbfModel.jl:793-795."""
        return self.jac(x, [ell])[0]


def build_model(sites: np.ndarray, values: np.ndarray, cfg: RbfConfig, shape: Optional[float] = None,
                fully_linear: bool = False) -> RbfModel:
    """A10: dense saddle-point system [Phi Pi; Pi' 0][w; lam] = [Y; 0], LU (`\\`), U5.

    Polynomial degree is raised to cpd_order-1 when too low (U4).  When there are fewer sites
    than polynomial basis functions the square system is singular; the minimum-norm solution
    is taken (U9, assumption).
    """
    sites = np.asarray(sites, dtype=np.float64)
    values = np.asarray(values, dtype=np.float64)
    N, n = sites.shape
    rf = get_radial_function(cfg, shape)
    deg = max(cfg.polynomial_degree, rf.cpd_order - 1)
    if deg > 2:
        raise NotImplementedError("polynomial tails of degree > 2 (thin plate splines of order >= 3, cubic exponents >= 7) are not built")
    p = poly_dim(n, deg)
    Phi = rf.phi(np.sqrt(((sites[:, None, :] - sites[None, :, :]) ** 2).sum(-1)))
    Pi = np.array([poly_basis(s, deg) for s in sites]).reshape(N, p)
    S = np.block([[Phi, Pi], [Pi.T, np.zeros((p, p))]])
    rhs = np.vstack((values, np.zeros((p, values.shape[1]))))
    if N >= p:
        coeff = np.linalg.solve(S, rhs)                                  # LAPACK gesv = LU, as `\`
        cond = float(np.linalg.cond(S))
    else:
        coeff = np.linalg.lstsq(S, rhs, rcond=None)[0]                   # U9
        cond = float("inf")
    return RbfModel(sites.copy(), coeff[:N], coeff[N:], rf, deg, fully_linear, cond)


def update_model(meta: RbfMeta, cfg: RbfConfig, db: ArrayDB, shape: Optional[float] = None) -> RbfModel:
    """RbfModel.jl:743-767."""
    ids = meta.collect_indices()
    sites = np.array([db.get_site(i) for i in ids])
    values = np.array([db.get_value(i) for i in ids])
    return build_model(sites, values, cfg, shape, meta.fully_linear)


# --------------------------------------------------------------------------------------
# A12  Armijo backtracking over surrogate values                       src/descent.jl:137-185
# --------------------------------------------------------------------------------------
def backtrack(model_eval, x, direction, step_size, omega, *, c=1e-6, shrink=0.75,
              min_stepsize=10 * EPS, max_loops=None, strict=True):
    if max_loops is None:
        max_loops = int(math.floor(math.log(min_stepsize) / math.log(shrink)))     # :62-66
    mx = model_eval(x)
    xp = x + step_size * direction
    mxp = model_eval(xp)
    i = 0
    while i < max_loops:
        if strict:
            ok = bool(np.all((mx - mxp) >= step_size * c * omega))
        else:
            ok = bool(np.max(mx) - np.max(mxp) >= step_size * c * omega)
        if ok:
            break
        if step_size <= min_stepsize:
            break
        step_size *= shrink
        xp = x + step_size * direction
        mxp = model_eval(xp)
        i += 1
    return xp, mxp, step_size * direction, i
