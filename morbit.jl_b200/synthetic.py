"""Seeded synthetic inputs of the BASELINE configs; pure NumPy so that the CUDA path and
the CPU oracle see identical bytes.  Nothing here is on the product's compute path."""
from __future__ import annotations

import numpy as np

def _first_primes(count: int):
    out, c = [], 2
    while len(out) < count:
        if all(c % q for q in out if q * q <= c):
            out.append(c)
        c += 1
    return out


_PRIMES = _first_primes(512)


def halton(num: int, dim: int, start: int = 1) -> np.ndarray:
    """First `num` points of the `dim`-dimensional Halton sequence (x0 generator, large_scale_benchmarks.jl:102-109 style)."""
    out = np.zeros((num, dim))
    idx = np.arange(start, start + num)
    for d in range(dim):
        base = _PRIMES[d]
        f, r, i = 1.0, np.zeros(num), idx.copy()
        while np.any(i > 0):
            f /= base
            r += f * (i % base)
            i //= base
        out[:, d] = r
    return out


def zdt1(X: np.ndarray) -> np.ndarray:
    X = np.clip(X, 0.0, 1.0)        # identity on the feasible box; the reference's wall steps can leave it by an ulp (sqrt domain)
    f1 = X[..., 0]
    g = 1.0 + 9.0 * np.mean(X[..., 1:], axis=-1)
    return np.stack([f1, g * (1.0 - np.sqrt(f1 / g))], axis=-1)


def zdt3(X: np.ndarray) -> np.ndarray:
    X = np.clip(X, 0.0, 1.0)
    f1 = X[..., 0]
    g = 1.0 + 9.0 * np.mean(X[..., 1:], axis=-1)
    h = 1.0 - np.sqrt(f1 / g) - (f1 / g) * np.sin(10.0 * np.pi * f1)
    return np.stack([f1, g * h], axis=-1)


def two_parabolas(X: np.ndarray) -> np.ndarray:
    """examples/example_two_parabolas.jl:38-39."""
    return np.stack([np.sum((X - 1.0) ** 2, axis=-1), np.sum((X + 1.0) ** 2, axis=-1)], axis=-1)


def multistart_batch(B: int, n: int = 30, n_db: int = 128, delta: float = 0.1, delta_max: float = 0.5,
                     theta_enlarge_2: float = 2.0, func=zdt3, first_instance: int = 0, local_fraction: float = 0.0) -> dict:
    """C3: instance b has iterate x = Halton point b and a database snapshot of n_db sites: the iterate (id 1) plus
    sites drawn uniformly in box 2 = [x - θ2 Δmax, x + θ2 Δmax] ∩ [0,1]^n (seed = instance id).  A fraction
    `local_fraction` of them is drawn in the trust region box of radius 2Δ instead (a database that already
    holds nearby evaluations)."""
    X0 = halton(first_instance + B, n)[first_instance:]
    sites = np.zeros((B, n_db, n))
    r2 = theta_enlarge_2 * delta_max
    for b in range(B):
        rng = np.random.default_rng(first_instance + b)
        x = X0[b]
        lo, hi = np.maximum(0.0, x - r2), np.minimum(1.0, x + r2)
        pts = lo + (hi - lo) * rng.random((n_db - 1, n))
        n_loc = int(round(local_fraction * (n_db - 1)))
        if n_loc:
            lo1, hi1 = np.maximum(0.0, x - 2 * delta), np.minimum(1.0, x + 2 * delta)
            pts[:n_loc] = lo1 + (hi1 - lo1) * rng.random((n_loc, n))
        sites[b, 0] = x
        sites[b, 1:] = pts
    values = func(sites)
    return dict(sites=sites, values=values, n_db=np.full(B, n_db, np.int32), x_index=np.ones(B, np.int32), x=X0.copy(),
                delta=np.full(B, delta), glb=np.zeros(n), gub=np.ones(n), flags_in=np.zeros((B, 2), np.int32),
                max_new=np.full(B, 2**31 - 1, np.int32), delta_max=delta_max)


def eval_sweep(N: int = 512, d: int = 50, k: int = 1, M: int = 10**6, seed: int = 0):
    """C5: N centres uniform in [0,1]^d (seed), values [sum x^2, sum sin x][:k]; M trial points uniform in
    [c_bar - 0.2, c_bar + 0.2] ∩ [0,1]^d (seed + 1)."""
    rng = np.random.default_rng(seed)
    centers = rng.random((N, d))
    vals = np.stack([np.sum(centers ** 2, axis=1), np.sum(np.sin(centers), axis=1)], axis=1)[:, :k]
    cbar = centers.mean(axis=0)
    lo, hi = np.maximum(0.0, cbar - 0.2), np.minimum(1.0, cbar + 0.2)
    rng2 = np.random.default_rng(seed + 1)
    X = lo + (hi - lo) * rng2.random((M, d))
    return centers, vals, X
