"""TEST INFRASTRUCTURE ONLY (CPU oracle; never imported by the product path).

Scalar restatement of one `optimize` run of Morbit.jl for the case the batched lock-step driver
(`morbit.jl_b200/lockstep.py`) covers: one RbfConfig group holding every objective, box constraints only (no
linear / nonlinear constraints => DummyFilter, theta_k = 0, no normal step), steepest descent with Armijo
backtracking, identity variable scaling (the problem is given in the scaled space: [0,1]^n or unbounded).

Follows, line by line:
  * initialize_data / init_surrogates     /root/reference/src/algorithm.jl:223-313, src/SurrogateContainer.jl:272-295
  * optimize loop                         src/algorithm.jl:919-958
  * iterate!                              src/algorithm.jl:615-917   (stopping tests :14-86, radius updates :151-196)
  * criticality_routine                   src/algorithm.jl:523-612   (quirk kept: the shrunken radius is a LOCAL variable,
                                          `update_surrogates!` is called with the unchanged iterate, :572-579)
  * get_criticality / compute_descent_step / _backtrack   src/descent.jl:187-241, 243-321, 150-185
  * update_surrogates! / improve_surrogates!              src/SurrogateContainer.jl:334-391
  * AlgorithmConfig defaults              src/AbstractConfigInterface.jl:14-95 -- the defaults are Float32 literals
                                          (MIN_PRECISION, e.g. 0.1f0) promoted to Float64 when they meet a Float64 iterate;
                                          `F32` below reproduces exactly those values.

PARITY UNPINNED (no Julia here, no golden vectors in the reference for this path; see oracle/rbf_oracle.py).
The LP of the steepest-descent direction has a degenerate optimal face on problems like ZDT (omega = 1 with many
optimal d); the reference's OSQP point, HiGHS' vertex and the GPU simplex' vertex are all optimal and all different.
`optimize(..., direction=callback)` therefore lets a test inject the direction the GPU path returned: the oracle
checks that it is feasible and optimal for its own Jacobian (oracle/descent_oracle.check_optimal) and continues
from it, so that everything else of the trajectory can be compared state by state.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np

from . import rbf_oracle as O
from . import descent_oracle as D

F32 = lambda v: float(np.float32(v))
_SQRT_EPS32 = float(np.sqrt(np.float32(np.finfo(np.float32).eps)))          # sqrt(eps(Float32)) in Float32
_OMEGA_TOL_REL = float(np.float32(10) * np.sqrt(np.float32(np.finfo(np.float32).eps)))

# return codes (src/Morbit.jl ITER/RET enums), iteration classifications
CONTINUE, MAX_ITER, BUDGET_EXHAUSTED, CRITICAL, TOLERANCE, INFEASIBLE = 0, 1, 2, 3, 4, 5
ACCEPTABLE, SUCCESSFULL, MODELIMPROVING, INACCEPTABLE, EARLY_EXIT = 0, 1, 2, 3, 4


@dataclass
class AlgoConfig:
    """AlgorithmConfig, src/ConfigImplementations.jl:13-85 with the defaults of src/AbstractConfigInterface.jl."""
    eps_crit: float = F32(0.001)
    gamma_crit: float = F32(0.51)
    max_critical_loops: int = 5
    delta_0: float = F32(0.1)
    delta_max: float = F32(0.5)
    max_evals: int = O.INT_MAX
    max_iter: int = 50
    f_tol_rel: float = _SQRT_EPS32
    x_tol_rel: float = _SQRT_EPS32
    f_tol_abs: float = -1.0
    x_tol_abs: float = -1.0
    omega_tol_rel: float = _OMEGA_TOL_REL
    delta_tol_rel: float = _SQRT_EPS32
    omega_tol_abs: float = -math.inf
    delta_tol_abs: float = _SQRT_EPS32
    stepnorm_tol_abs: float = 0.0
    strict_acceptance_test: bool = True
    nu_success: float = F32(0.2)
    nu_accept: float = 0.0
    mu: float = F32(2e3)
    beta: float = F32(1e3)
    gamma_grow: float = 2.0
    gamma_shrink: float = 0.75
    gamma_shrink_much: float = F32(0.51)
    # SteepestDescentConfig, src/descent.jl:51-73
    strict_backtracking: bool = True
    armijo_const_rhs: float = 1e-6
    armijo_const_shrink: float = 0.75
    min_stepsize: float = 10 * O.EPS
    normalize: bool = True
    isapprox_rtol: float = O.ISAPPROX_RTOL_F32      # `Δ ≈ Δ_max` with a Float32 delta_max (RbfModel.jl:588, AbstractConfigInterface.jl:31)


@dataclass
class IterRecord:
    """State after one iterate! call (what the lock-step driver records too)."""
    iter_counter: int
    ret_code: int
    it_stat: int
    x: np.ndarray
    fx: np.ndarray
    x_index: int
    delta: float
    n_db: int
    num_evals: int
    omega: float = math.nan
    rho: float = math.nan
    steplength: float = math.nan
    n_crit_loops: int = 0
    training_ids: List[int] = field(default_factory=list)
    fully_linear: bool = False
    knife: bool = False          # some threshold decision on the way here was closer than rounding accuracy
    wall_tie: bool = False       # a model update of THIS iteration placed a site by comparing two equal wall distances (sign = rounding noise)


class Run:
    """One `optimize(mop, x0)` (src/algorithm.jl:919-958)."""

    def __init__(self, func: Callable, x0, glb, gub, cfg: O.RbfConfig, ac: Optional[AlgoConfig] = None,
                 direction: Optional[Callable] = None):
        self.func, self.cfg, self.ac = func, cfg, ac or AlgoConfig()
        self.glb, self.gub = np.asarray(glb, np.float64), np.asarray(gub, np.float64)
        self.direction = direction
        assert self.ac.delta_max <= 1.0, "compute_descent_step's delta > 1 branch (descent.jl:276-311) is not restated"
        # initialize_data, :223-313
        x = np.clip(np.asarray(x0, np.float64), self.glb, self.gub)          # _project_into_box, :256-258
        self.db = O.ArrayDB()
        self.fx = np.atleast_1d(np.asarray(func(x), np.float64))
        self.num_evals = 1
        self.x, self.x_index = x, self.db.new_result(x, self.fx)
        self.delta = float(self.ac.delta_0)
        self.meta = O.RbfMeta(signature=cfg.signature())
        self.model = None
        self.iter_counter, self.it_stat, self.ret_code = 1, ACCEPTABLE, CONTINUE
        self.records: List[IterRecord] = []
        self.knife = False
        # a round-3 / improvement step in a box that is symmetric about the iterate (unbounded problems): `intersect_box(:absmax)`
        # compared two equal wall distances, so the SIGN of the new site follows the last bits of x and of the direction
        self.wall_tie = False
        self._update(True)                                                   # init_surrogates: prepare_init_model => ensure_fully_linear

    # ------------------------------------------------------------------ surrogates
    def _eval_missing(self):
        self.num_evals += self.db.eval_missing(self.func)

    def _update(self, ensure_fully_linear: bool):
        """update_surrogates!, SurrogateContainer.jl:339-390 (one group)."""
        tf, t4 = O.FilterTrace(), O.Round4Trace()
        O.WALL_TIES.clear()
        self.meta = O.prepare_update_model(self.meta, self.cfg, self.db, self.x, self.x_index, self.delta, self.ac.delta_max,
                                           self.glb, self.gub, ensure_fully_linear=ensure_fully_linear,
                                           num_objf_evals=self.num_evals, algo_max_evals=self.ac.max_evals, trace=tf, trace4=t4,
                                           isapprox_rtol=self.ac.isapprox_rtol)
        # decisions taken by less than rounding accuracy (pivot tests of the filter, tau^2 of round 4): a second implementation
        # may legitimately decide the other way, so a comparison has to stop at this state
        self.knife = self.knife or tf.knife_edge() or any(abs(v) < 1e-12 for v in t4.tau2)
        self.wall_tie = self.wall_tie or bool(O.WALL_TIES)
        self._eval_missing()
        self.model = O.update_model(self.meta, self.cfg, self.db)

    def _improve(self):
        O.WALL_TIES.clear()
        self.meta = O.prepare_improve_model(self.meta, self.cfg, self.db, self.x, self.delta, self.glb, self.gub)
        self.wall_tie = self.wall_tie or bool(O.WALL_TIES)
        self._eval_missing()
        self.model = O.update_model(self.meta, self.cfg, self.db)

    def _criticality(self):
        """get_criticality(::SteepestDescentConfig), descent.jl:187-241 (no constraints: x_n = x)."""
        J = self.model.jac(self.x)
        if self.direction is not None:
            d, omega = self.direction(self, J)
            if not self.knife:                                     # after a knife-edge decision the two states may differ legitimately
                D.check_optimal(self.x, J, self.glb, self.gub, d, omega, self.ac.normalize)
            return float(omega), np.asarray(d, np.float64)
        d, omega = D.lp_highs(self.x, J, self.glb, self.gub, self.ac.normalize)
        return float(omega), d

    def _budget_okay(self) -> bool:
        return self.num_evals < min(self.cfg.max_evals, self.ac.max_evals)   # VecFun.jl:318-320, algorithm.jl:5-11

    # ------------------------------------------------------------------ criticality_routine, algorithm.jl:523-612
    def _criticality_routine(self, omega, d):
        ac = self.ac
        beta = max(ac.beta, ac.mu)
        do_loops, loops, exit_critical = True, 0, False
        if not self.meta.fully_linear:
            self._update(True)
            omega, d = self._criticality()
            do_loops = self.delta > ac.mu * omega if self.meta.fully_linear else False
        if do_loops:
            delta = delta_0 = self.delta
            while delta > ac.mu * omega:
                if loops >= ac.max_critical_loops or not self._budget_okay():
                    exit_critical = True
                    break
                delta = ac.gamma_crit * delta
                self._update(True)                                            # quirk: the iterate still carries the old radius
                omega, d = self._criticality()
                loops += 1
                if (delta <= ac.delta_tol_abs or (omega <= ac.omega_tol_rel and delta <= ac.delta_tol_rel)
                        or omega <= ac.omega_tol_abs):
                    exit_critical = True
                    break
                if not self.meta.fully_linear:
                    exit_critical = True
                    break
            self.delta = min(delta_0, max(beta * omega, delta))               # :603
        return exit_critical, omega, d, loops

    # ------------------------------------------------------------------ iterate!, algorithm.jl:615-917
    def iterate(self):
        ac = self.ac
        rec = lambda **kw: self._record(**kw)
        if self.iter_counter > ac.max_iter:
            return rec(ret=MAX_ITER, stat=EARLY_EXIT)
        if not self._budget_okay():
            return rec(ret=BUDGET_EXHAUSTED, stat=EARLY_EXIT)
        if self.delta <= ac.delta_tol_abs:
            return rec(ret=TOLERANCE, stat=EARLY_EXIT)
        beta = max(ac.beta, ac.mu)
        if self.iter_counter > 1:
            if self.it_stat == MODELIMPROVING:
                self._improve()
            else:
                self._update(False)
        omega, d = self._criticality()
        if (omega <= ac.omega_tol_rel and self.delta <= ac.delta_tol_rel) or omega <= ac.omega_tol_abs:
            return rec(ret=CRITICAL, stat=EARLY_EXIT, omega=omega)
        loops = 0
        fl = self.meta.fully_linear
        if omega <= ac.eps_crit and (not fl or self.delta > ac.mu * omega):
            exit_critical, omega, d, loops = self._criticality_routine(omega, d)
            if exit_critical:
                return rec(ret=CRITICAL, stat=EARLY_EXIT, omega=omega, loops=loops)
        # compute_descent_step, descent.jl:243-321 (delta <= 1 branch)
        x = self.x
        norm_d = float(np.max(np.abs(d)))
        with np.errstate(divide="ignore", invalid="ignore"):
            sigma = min(self.delta / norm_d, 1.0) if norm_d > 0 else 1.0        # Δ/0 = Inf in Julia => min(Inf, 1) = 1
        if sigma > ac.min_stepsize:
            x_trial, mx_trial_bt, step, _ = O.backtrack(self.model.eval, x, d, sigma, omega, c=ac.armijo_const_rhs,
                                                        shrink=ac.armijo_const_shrink, min_stepsize=ac.min_stepsize,
                                                        strict=ac.strict_backtracking)
            steplength = float(np.max(np.abs(step)))
        else:
            omega, x_trial, steplength = 0.0, x.copy(), 0.0
        fx_trial = np.atleast_1d(np.asarray(self.func(x_trial), np.float64))    # :760
        self.num_evals += 1
        new_index = self.db.new_result(x_trial, fx_trial)                       # put_eval_result_into_db!, :764
        mx, mx_trial = self.model.eval(x), self.model.eval(x_trial)             # :766-767
        steplength = float(np.max(np.abs(x - x_trial)))                         # :773 (recomputed from the points)
        # acceptance, :793-866 (DummyFilter: acceptable, theta_k = 0)
        if ac.strict_acceptance_test:
            denom = mx - mx_trial
            rho = math.nan if np.any(denom == 0) else float(np.min((self.fx - fx_trial) / denom))
            good = bool(np.all(denom >= 0.0))
        else:
            denom = float(np.max(mx) - np.max(mx_trial))
            with np.errstate(divide="ignore", invalid="ignore"):
                rho = float(np.float64(np.max(self.fx) - np.max(fx_trial)) / np.float64(denom))
            good = denom >= 0.0
        rho = -math.inf if math.isnan(rho) else rho
        accept, stat, upd = True, ACCEPTABLE, "leave"
        if good:
            if rho >= ac.nu_success:
                stat = SUCCESSFULL
                if self.delta < beta * omega:
                    upd = "grow"
            elif self.meta.fully_linear:
                if rho >= ac.nu_accept:
                    stat, upd = ACCEPTABLE, "shrink"
                else:
                    accept, stat, upd = False, INACCEPTABLE, "shrink_much"
            else:
                accept, stat, upd = False, MODELIMPROVING, "leave"
        else:
            FILTER_ADD = 5
            stat, upd = FILTER_ADD, ("grow" if rho >= ac.nu_success else "leave")
        if not accept and steplength <= ac.stepnorm_tol_abs:
            return rec(ret=TOLERANCE, stat=stat, omega=omega, rho=rho, steplength=steplength, loops=loops)
        if upd == "grow":
            self.delta = min(ac.delta_max, ac.gamma_grow * self.delta)
        elif upd == "shrink":
            self.delta = self.delta * ac.gamma_shrink
        elif upd == "shrink_much":
            self.delta = self.delta * ac.gamma_shrink_much
        ret = CONTINUE
        if accept:
            x_old, fx_old = self.x, self.fx
            self.x, self.fx, self.x_index = x_trial, fx_trial, new_index
            dx, df = float(np.max(np.abs(x_old - x_trial))), float(np.max(np.abs(fx_old - fx_trial)))
            if (dx <= ac.x_tol_rel * float(np.max(np.abs(x_old))) or dx <= ac.x_tol_abs
                    or df <= ac.f_tol_rel * float(np.max(np.abs(fx_old))) or df <= ac.f_tol_abs):
                ret = TOLERANCE
        return rec(ret=ret, stat=stat, omega=omega, rho=rho, steplength=steplength, loops=loops)

    def _record(self, ret, stat, omega=math.nan, rho=math.nan, steplength=math.nan, loops=0):
        self.ret_code, self.it_stat = ret, stat
        r = IterRecord(self.iter_counter, ret, stat, self.x.copy(), self.fx.copy(), self.x_index, self.delta, self.db.num_entries,
                       self.num_evals, omega, rho, steplength, loops, list(self.meta.collect_indices()), self.meta.fully_linear, self.knife, self.wall_tie)
        self.wall_tie = False
        self.records.append(r)
        self.iter_counter += 1
        return r

    def run(self):
        while self.ret_code == CONTINUE:
            self.iterate()
        return self.x, self.fx, self.ret_code


def optimize(func, x0, glb, gub, cfg: O.RbfConfig, ac: Optional[AlgoConfig] = None, direction=None) -> Run:
    r = Run(func, x0, glb, gub, cfg, ac, direction)
    r.run()
    return r
