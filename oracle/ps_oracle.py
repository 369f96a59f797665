"""TEST INFRASTRUCTURE (oracle): CPU restatement of the Pascoletti-Serafini inner solves of Morbit.jl.

Reference path: src/descent.jl:369-387 (`_min_component`), :404-412 (`compute_local_ideal_point`), :435-476 (constraint and objective
handles of the PS problem), :478-500 (`_ps_optimization`), :503-581 (`get_criticality(::PascolettiSerafiniConfig)`).

The arithmetic of the solver itself lives in an un-vendored dependency: NLopt.jl (`Project.toml` compat "0.6", no Manifest) ->
libnlopt 2.7, algorithm `:GN_ISRES` = Runarsson & Yao, "Search biases in constrained evolutionary optimization" (IEEE Trans. SMC-C 35,
2005): a (mu, lambda) evolution strategy with stochastic ranking, lambda = 20 (n + 1), mu = lambda / 7, log-normal self-adaptation
(phi = 1), smoothing alpha = 0.2, differential variation gamma = 0.85 for the mu - 1 best, ranking probability pf = 0.45.  NLopt
drives it with its own Mersenne-Twister stream, so NLopt's iterates cannot be reproduced by anything but NLopt: **parity unpinned**
for this row -- what is checked is (i) this restatement against the CUDA path generation by generation (same counter-based random
numbers), and (ii) the solution quality of both against an independent multi-start SQP solve (scipy SLSQP) of the same problem.

Two problem kinds (matching mrbf_ps_solve):
  ideal point:  min m_l(xi)  over the box, constraint surrogates c(xi) <= 0
  PS:           min tau  s.t.  m_l(xi) - m_l(x) - tau r_l <= 0, -1 <= tau <= 0, box, c(xi) <= 0 -- tau eliminated:
                tau(xi) = clamp(max_l (m_l(xi) - m_l(x)) / r_l, -1, 0)
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import numpy as np

ALPHA, GAMMA, PF, RETRY = 0.2, 0.85, 0.45, 10
_M64 = (1 << 64) - 1


def splitmix64(z: int) -> int:
    z = (z + 0x9E3779B97F4A7C15) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def key(seed: int, b: int, gen: int, i: int, j: int, stream: int) -> int:
    h = splitmix64(seed ^ ((0xD1B54A32D192ED03 * (b + 1)) & _M64))
    h = splitmix64(h ^ ((((gen & 0xFFFFFFFF) << 32) | (i & 0xFFFFFFFF)) & _M64))
    return splitmix64(h ^ ((((j & 0xFFFFFFFF) << 8) | stream) & _M64))


def u01(h: int) -> float:
    return (float(h >> 11) + 0.5) * (1.0 / 9007199254740992.0)


def normal01(h: int) -> float:
    u1, u2 = u01(h), u01(splitmix64(h))
    return math.sqrt(-2.0 * math.log(u1)) * math.cos(2.0 * math.pi * u2)


def fitness_penalty(Y: np.ndarray, n_obj: int, objective: int, mx: Optional[np.ndarray], direction: Optional[np.ndarray],
                    first_is_start: bool = False):
    """Y: lambda x k surrogate values -> (f, phi)."""
    lam, k = Y.shape
    if direction is not None:
        t = np.max((Y[:, :n_obj] - mx[None, :n_obj]) / direction[None, :n_obj], axis=1)
        # t <= 0: feasible at tau = clamp(t) by construction; t > 0: the violation left at tau = 0.  The start point (individual 0 of
        # generation 0) is feasible with tau = 0 by definition (m(x) - m(x) - 0 r = 0).
        pen = np.where(t > 0.0, (np.maximum(Y[:, :n_obj] - mx[None, :n_obj], 0.0) ** 2).sum(1), 0.0)
        if first_is_start:
            pen[0] = 0.0
        f = np.minimum(np.maximum(t, -1.0), 0.0)
    else:
        f = Y[:, objective].copy()
        pen = np.zeros(lam)
    if k > n_obj:
        pen = pen + (np.maximum(Y[:, n_obj:], 0.0) ** 2).sum(1)
    bad = ~(np.isfinite(f) & np.isfinite(pen))
    f[bad] = np.inf; pen[bad] = np.inf
    return f, pen


def stochastic_rank(f, phi, seed, b, gen):
    """Odd-even transposition form of Runarsson & Yao's stochastic ranking: lambda phases of disjoint neighbour comparisons."""
    lam = len(f)
    idx = list(range(lam))
    for ph in range(lam):
        for a in range(ph & 1, lam - 1, 2):
            ia, ib = idx[a], idx[a + 1]
            u = u01(key(seed, b, gen, ph, a, 1))
            if (phi[ia] == 0.0 and phi[ib] == 0.0) or u < PF:
                swap = f[ia] > f[ib]
            else:
                swap = phi[ia] > phi[ib]
            if swap:
                idx[a], idx[a + 1] = ib, ia
    return idx


def ps_solve(model_eval: Callable[[np.ndarray], np.ndarray], x0, lb, ub, mx=None, direction=None, n_obj: Optional[int] = None,
             objective: int = -1, population: int = -1, max_evals: int = -1, seed: int = 0, b: int = 0, trace: Optional[list] = None):
    """One instance.  model_eval: (M x n) -> (M x k).  Returns (f_min, x_min, y_min, found, evals)."""
    x0 = np.asarray(x0, dtype=np.float64); lb = np.asarray(lb, dtype=np.float64); ub = np.asarray(ub, dtype=np.float64)
    n = len(x0)
    lam = population if population > 0 else 20 * (n + 1)
    lam = max(lam, 8)
    evals = max_evals if max_evals > 0 else 500 * (n + 1)
    gens = max(evals // lam, 1)
    mu = (lam + 6) // 7
    rs = 1.0 / math.sqrt(n)
    X = np.empty((lam, n)); S = np.empty((lam, n))
    for i in range(lam):
        for j in range(n):
            w = ub[j] - lb[j]
            v = x0[j] if i == 0 else lb[j] + u01(key(seed, b, 0, i, j, 0)) * w
            X[i, j] = min(max(v, lb[j]), ub[j]); S[i, j] = w * rs
    k = None
    best_f, best_x, best_y, found = math.inf, x0.copy(), None, 0
    taup, tau = 1.0 / math.sqrt(2.0 * n), 1.0 / math.sqrt(2.0 * math.sqrt(n))
    for g in range(gens + 1):
        Y = np.asarray(model_eval(X), dtype=np.float64)
        k = Y.shape[1]
        no = k if n_obj is None else n_obj
        f, phi = fitness_penalty(Y, no, objective, mx, direction, g == 0)
        feas = np.where((phi == 0.0) & np.isfinite(f))[0]
        if len(feas):
            ib = feas[np.argmin(f[feas])]                  # first minimiser
            if f[ib] < best_f or not found:
                best_f, best_x, best_y, found = float(f[ib]), X[ib].copy(), Y[ib].copy(), 1
        if trace is not None:
            trace.append(dict(X=X.copy(), f=f.copy(), phi=phi.copy(), best_f=best_f))
        if g == gens:
            break
        rank = stochastic_rank(f, phi, seed, b, g)
        Xn = np.empty_like(X); Sn = np.empty_like(S)
        gn = g + 1
        for i in range(lam):
            par = rank[i % mu]
            gi = normal01(key(seed, b, gn, i, n, 2))
            for j in range(n):
                xp, sp, lo, hi = X[par, j], S[par, j], lb[j], ub[j]
                xv, sv, done = xp, sp, False
                if i < mu - 1:
                    cand = xp + GAMMA * (X[rank[0], j] - X[rank[i + 1], j])
                    if lo <= cand <= hi:
                        xv, done = cand, True
                if not done:
                    s1 = sp * math.exp(taup * gi + tau * normal01(key(seed, b, gn, i, j, 3)))
                    s1 = min(s1, (hi - lo) * rs)
                    for t in range(RETRY):
                        cand = xp + s1 * normal01(key(seed, b, gn, i, j, 4 + t))
                        if lo <= cand <= hi:
                            xv = cand
                            break
                    sv = sp + ALPHA * (s1 - sp)
                Xn[i, j] = xv; Sn[i, j] = sv
        X, S = Xn, Sn
    if best_y is None:
        best_y = np.full(k, np.nan)
    return best_f, best_x, best_y, found, (gens + 1) * lam


def reference_optimum(model_eval, model_jac, x0, lb, ub, mx=None, direction=None, n_obj: Optional[int] = None, objective: int = -1,
                      starts: int = 12, seed: int = 0):
    """Independent check of solution quality: multi-start SLSQP on the smooth formulation the reference hands to NLopt (variables
    (tau, xi) for PS).  Returns the best objective value found."""
    from scipy.optimize import minimize
    x0 = np.asarray(x0, dtype=np.float64); lb = np.asarray(lb, dtype=np.float64); ub = np.asarray(ub, dtype=np.float64)
    n = len(x0)
    rng = np.random.default_rng(seed)
    k = np.asarray(model_eval(x0[None]))[0].shape[0]
    no = k if n_obj is None else n_obj
    best = math.inf
    for s in range(starts):
        xs = x0 if s == 0 else lb + rng.random(n) * (ub - lb)
        if direction is not None:
            def fun(z): return z[0]
            def jac(z): g = np.zeros(n + 1); g[0] = 1.0; return g
            cons = [dict(type="ineq", fun=(lambda z, l=l: -(model_eval(z[None, 1:])[0][l] - mx[l] - z[0] * direction[l])),
                         jac=(lambda z, l=l: -np.concatenate(([-direction[l]], model_jac(z[1:])[l])))) for l in range(no)]
            cons += [dict(type="ineq", fun=(lambda z, l=l: -model_eval(z[None, 1:])[0][l]),
                          jac=(lambda z, l=l: -np.concatenate(([0.0], model_jac(z[1:])[l])))) for l in range(no, k)]
            t0 = max(-1.0, min(0.0, float(np.max((model_eval(xs[None])[0][:no] - mx[:no]) / direction[:no]))))
            res = minimize(fun, np.concatenate(([t0], xs)), jac=jac, constraints=cons, method="SLSQP",
                           bounds=[(-1.0, 0.0)] + list(zip(lb, ub)), options=dict(maxiter=300, ftol=1e-12))
            ok = res.success and all(c["fun"](res.x) >= -1e-7 for c in cons)
        else:
            cons = [dict(type="ineq", fun=(lambda z, l=l: -model_eval(z[None])[0][l]), jac=(lambda z, l=l: -model_jac(z)[l]))
                    for l in range(no, k)]
            res = minimize(lambda z: model_eval(z[None])[0][objective], xs, jac=lambda z: model_jac(z)[objective], constraints=cons,
                           method="SLSQP", bounds=list(zip(lb, ub)), options=dict(maxiter=300, ftol=1e-12))
            ok = res.success and all(c["fun"](res.x) >= -1e-7 for c in cons)
        if ok and res.fun < best:
            best = float(res.fun)
    return best
