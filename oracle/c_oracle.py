"""ctypes front-end of oracle/rbf_oracle.c (TEST INFRASTRUCTURE ONLY, see that file's header).

The shared object is compiled with -march=native, so it is rebuilt whenever the host CPU
differs from the one it was built on (the build box and the GPU box are different machines).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import math
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librbf_oracle.so")
_TAG = os.path.join(_HERE, "librbf_oracle.host")

KERNEL_IDS = {"cubic": 0, "inv_multiquadric": 1, "multiquadric": 2, "thin_plate_spline": 3, "gaussian": 4}


def _host_tag() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        model = [l for l in txt.splitlines() if l.startswith(("model name", "flags"))][:2]
        return hashlib.sha1("".join(model).encode()).hexdigest()
    except OSError:
        return "unknown"


def build(force: bool = False) -> str:
    tag = _host_tag()
    src = os.path.join(_HERE, "rbf_oracle.c")
    fresh = (os.path.exists(_SO) and os.path.exists(_TAG) and open(_TAG).read().strip() == tag
             and os.path.getmtime(_SO) >= os.path.getmtime(src))
    if force or not fresh:
        subprocess.run(["make", "-C", _HERE, "-B", "librbf_oracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        with open(_TAG, "w") as f:
            f.write(tag)
    return _SO


class _Cfg(C.Structure):
    _fields_ = [("kernel", C.c_int32), ("poly_degree", C.c_int32), ("alpha", C.c_double), ("beta", C.c_double),
                ("theta_enlarge_1", C.c_double), ("theta_enlarge_2", C.c_double), ("theta_pivot", C.c_double),
                ("theta_pivot_cholesky", C.c_double), ("max_model_points", C.c_int32),
                ("optimized_sampling", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_intersect_box_absmax.restype = C.c_double
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def make_cfg(cfg, shape: Optional[float] = None) -> _Cfg:
    """cfg: any object with RbfConfig's fields (oracle.rbf_oracle.RbfConfig or the product's)."""
    sp = cfg.shape_parameter if shape is None else shape
    nan = isinstance(sp, float) and math.isnan(sp)
    k = cfg.kernel
    alpha, beta = 1.0, 0.0
    if k == "gaussian":
        alpha = 1.0 if nan else float(sp)
    elif k in ("multiquadric", "inv_multiquadric"):
        alpha, beta = (1.0 if nan else float(sp)), 0.5
    elif k == "cubic":
        beta = 3.0 if nan else float(int(sp))
    elif k == "thin_plate_spline":
        beta = 2.0 if nan else float(int(sp))
    return _Cfg(KERNEL_IDS[k], cfg.polynomial_degree, alpha, beta, cfg.theta_enlarge_1, cfg.theta_enlarge_2,
                cfg.theta_pivot, cfg.theta_pivot_cholesky, cfg.max_model_points, int(bool(cfg.optimized_sampling)))


def max_points(cfg, n: int) -> int:
    return ((n + 1) * (n + 2)) // 2 if cfg.max_model_points <= 0 else cfg.max_model_points


def max_threads() -> int:
    return int(lib().orc_max_threads())


@dataclass
class SelectResult:
    r1: np.ndarray; n_r1: np.ndarray
    r2: np.ndarray; n_r2: np.ndarray
    r3_sites: np.ndarray; n_r3: np.ndarray
    r4: np.ndarray; n_r4: np.ndarray
    dirs: np.ndarray; n_dirs: np.ndarray
    fully_linear: np.ndarray; rebuilt: np.ndarray
    margins: np.ndarray


def select_points_batched(cfg, sites, x_index, x, delta, delta_max, glb, gub, ensure_fully_linear, force_rebuild,
                          max_new, nthreads: int = 1) -> SelectResult:
    """sites: (B, n_db, n); x: (B, n); x_index/delta/flags/max_new: (B,)."""
    sites = np.ascontiguousarray(sites, dtype=np.float64)
    B, n_db, n = sites.shape
    x = np.ascontiguousarray(x, dtype=np.float64)
    x_index = np.ascontiguousarray(x_index, dtype=np.int32)
    delta = np.ascontiguousarray(delta, dtype=np.float64)
    glb = np.ascontiguousarray(np.broadcast_to(glb, (n,)), dtype=np.float64)
    gub = np.ascontiguousarray(np.broadcast_to(gub, (n,)), dtype=np.float64)
    flags_in = np.ascontiguousarray(np.stack([np.broadcast_to(ensure_fully_linear, (B,)),
                                              np.broadcast_to(force_rebuild, (B,))], axis=1), dtype=np.int32)
    max_new = np.ascontiguousarray(np.broadcast_to(max_new, (B,)), dtype=np.int32)
    mp = max_points(cfg, n)
    r1 = np.zeros((B, n), np.int32); n_r1 = np.zeros(B, np.int32)
    r2 = np.zeros((B, n), np.int32); n_r2 = np.zeros(B, np.int32)
    r3 = np.zeros((B, n, n), np.float64); n_r3 = np.zeros(B, np.int32)
    r4 = np.zeros((B, mp), np.int32); n_r4 = np.zeros(B, np.int32)
    dirs = np.zeros((B, n, n), np.float64); n_dirs = np.zeros(B, np.int32)
    flags_out = np.zeros((B, 2), np.int32)
    margins = np.zeros((B, 2), np.float64)
    ccfg = make_cfg(cfg)
    lib().orc_select_points_batched(C.byref(ccfg), B, n, n_db, _dp(sites), _ip(x_index), _dp(x), _dp(delta),
                                    C.c_double(delta_max), _dp(glb), _dp(gub), _ip(flags_in), _ip(max_new),
                                    _ip(r1), _ip(n_r1), _ip(r2), _ip(n_r2), _dp(r3), _ip(n_r3), mp, _ip(r4), _ip(n_r4),
                                    _dp(dirs), _ip(n_dirs), _ip(flags_out), _dp(margins), nthreads)
    return SelectResult(r1, n_r1, r2, n_r2, r3, n_r3, r4, n_r4, dirs, n_dirs,
                        flags_out[:, 0].astype(bool), flags_out[:, 1].astype(bool), margins)


FUNC_IDS = {None: 0, "zdt3": 1, "zdt1": 2, "two_parabolas": 3}


def set_isapprox_rtol(rtol: float) -> None:
    """rtol of `isapprox(delta, delta_max)` (RbfModel.jl:588): sqrt(eps(Float64)) by default, sqrt(eps(Float32)) when the
    algorithm config holds Float32 radii (the reference's default config)."""
    lib().orc_set_isapprox_rtol(C.c_double(rtol))


def select_and_build_batched(cfg, sites, values, x_index, x, delta, delta_max, glb, gub, ensure_fully_linear=False,
                             force_rebuild=False, max_new=2**31 - 1, func: Optional[str] = None, nthreads: int = 1):
    """One whole model build per instance and thread inside the C library (rounds 1-4, training-set gather from the database
    arrays, objective values of the new round-3 sites, saddle solve): the CPU arm of bench.py.  Returns (N, ids, w, lam, status)."""
    sites = np.ascontiguousarray(sites, dtype=np.float64); values = np.ascontiguousarray(values, dtype=np.float64)
    B, n_db, n = sites.shape
    k = values.shape[2]
    x = np.ascontiguousarray(x, dtype=np.float64); x_index = np.ascontiguousarray(x_index, dtype=np.int32)
    delta = np.ascontiguousarray(delta, dtype=np.float64)
    glb = np.ascontiguousarray(np.broadcast_to(glb, (n,)), dtype=np.float64)
    gub = np.ascontiguousarray(np.broadcast_to(gub, (n,)), dtype=np.float64)
    flags_in = np.ascontiguousarray(np.stack([np.broadcast_to(ensure_fully_linear, (B,)),
                                              np.broadcast_to(force_rebuild, (B,))], axis=1), dtype=np.int32)
    max_new = np.ascontiguousarray(np.broadcast_to(max_new, (B,)), dtype=np.int32)
    ccfg = make_cfg(cfg)
    p = _poly_dim_c(ccfg, n)
    ts = max(n + 1, min(max_points(cfg, n), n + n_db))
    N = np.zeros(B, np.int32); ids = np.zeros((B, ts), np.int32)
    w = np.zeros((B, ts, k)); lam = np.zeros((B, p, k)); status = np.zeros(B, np.int32)
    lib().orc_select_and_build_batched(C.byref(ccfg), B, n, k, n_db, _dp(sites), _dp(values), _ip(x_index), _dp(x), _dp(delta),
                                       C.c_double(delta_max), _dp(glb), _dp(gub), _ip(flags_in), _ip(max_new), FUNC_IDS[func],
                                       ts, _ip(N), _ip(ids), ts, _dp(w), _dp(lam), _ip(status), nthreads)
    return N, ids, w, lam, status


def round4(cfg, sites, lb2, ub2, found):
    sites = np.ascontiguousarray(sites, dtype=np.float64)
    n_db, n = sites.shape
    found = np.ascontiguousarray(found, dtype=np.int32)
    r4 = np.zeros(max(max_points(cfg, n), 1), np.int32)
    margins = np.zeros(2)
    ccfg = make_cfg(cfg)
    lb2 = np.ascontiguousarray(lb2, dtype=np.float64); ub2 = np.ascontiguousarray(ub2, dtype=np.float64)
    k = lib().orc_round4(C.byref(ccfg), n, n_db, _dp(sites), _dp(lb2), _dp(ub2), _ip(found), len(found), _ip(r4),
                         _dp(margins))
    return r4[:k].copy(), margins[1]


def build_batched(cfg, sites, values, N, nthreads: int = 1, shape: Optional[float] = None):
    """sites (B, Ns, n), values (B, Ns, k), N (B,) -> w (B, Ns, k), lam (B, p, k), status (B,)."""
    sites = np.ascontiguousarray(sites, dtype=np.float64)
    values = np.ascontiguousarray(values, dtype=np.float64)
    B, Ns, n = sites.shape
    k = values.shape[2]
    N = np.ascontiguousarray(np.broadcast_to(N, (B,)), dtype=np.int32)
    ccfg = make_cfg(cfg, shape)
    p = _poly_dim_c(ccfg, n)
    w = np.zeros((B, Ns, k)); lam = np.zeros((B, p, k)); status = np.zeros(B, np.int32)
    lib().orc_build_batched(C.byref(ccfg), B, n, k, Ns, _ip(N), _dp(sites), _dp(values), _dp(w), _dp(lam), _ip(status),
                            nthreads)
    return w, lam, status


def _poly_dim_c(ccfg: _Cfg, n: int) -> int:
    cpd = {0: math.ceil(ccfg.beta / 2), 2: math.ceil(ccfg.beta), 3: int(ccfg.beta) + 1}.get(ccfg.kernel, 0)
    deg = max(ccfg.poly_degree, cpd - 1)
    return 0 if deg < 0 else (1 if deg == 0 else n + 1)


def eval_points(cfg, centers, w, lam, X, nthreads: int = 1, shape: Optional[float] = None):
    centers = np.ascontiguousarray(centers, dtype=np.float64); w = np.ascontiguousarray(w, dtype=np.float64)
    lam = np.ascontiguousarray(lam, dtype=np.float64); X = np.ascontiguousarray(X, dtype=np.float64)
    N, n = centers.shape; k = w.shape[1]; M = X.shape[0]
    Y = np.zeros((M, k))
    ccfg = make_cfg(cfg, shape)
    lib().orc_eval(C.byref(ccfg), n, k, N, _dp(centers), _dp(w), _dp(lam), C.c_long(M), _dp(X), _dp(Y), nthreads)
    return Y


def jac_points(cfg, centers, w, lam, X, nthreads: int = 1, shape: Optional[float] = None):
    centers = np.ascontiguousarray(centers, dtype=np.float64); w = np.ascontiguousarray(w, dtype=np.float64)
    lam = np.ascontiguousarray(lam, dtype=np.float64); X = np.ascontiguousarray(X, dtype=np.float64)
    N, n = centers.shape; k = w.shape[1]; M = X.shape[0]
    J = np.zeros((M, k, n))
    ccfg = make_cfg(cfg, shape)
    lib().orc_jac(C.byref(ccfg), n, k, N, _dp(centers), _dp(w), _dp(lam), C.c_long(M), _dp(X), _dp(J), nthreads)
    return J


def intersect_box_absmax(x, d, lb, ub) -> float:
    x, d, lb, ub = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, d, lb, ub))
    return float(lib().orc_intersect_box_absmax(len(x), _dp(x), _dp(d), _dp(lb), _dp(ub)))
