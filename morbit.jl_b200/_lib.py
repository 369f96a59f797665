"""ctypes binding of include/morbit_rbf.h.  There is no CPU fallback: if the CUDA library is
missing or no GPU is visible, every compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmorbit_rbf.so")

KERNEL_IDS = {"cubic": 0, "inv_multiquadric": 1, "multiquadric": 2, "thin_plate_spline": 3, "gaussian": 4}

MRBF_OK, MRBF_EINVAL, MRBF_ECUDA, MRBF_ENOMEM, MRBF_EUNSUPPORTED, MRBF_ENUMERIC = 0, -1, -2, -3, -4, -5
_CODES = {-1: "MRBF_EINVAL", -2: "MRBF_ECUDA", -3: "MRBF_ENOMEM", -4: "MRBF_EUNSUPPORTED", -5: "MRBF_ENUMERIC"}


class MrbfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{_CODES.get(code, code)}: {msg}")
        self.code = code


class MrbfCfg(C.Structure):
    """struct mrbf_cfg (include/morbit_rbf.h)."""
    _fields_ = [("kernel", C.c_int32), ("polynomial_degree", C.c_int32), ("shape_parameter", C.c_double),
                ("theta_enlarge_1", C.c_double), ("theta_enlarge_2", C.c_double), ("theta_pivot", C.c_double),
                ("theta_pivot_cholesky", C.c_double), ("max_model_points", C.c_int32), ("use_max_points", C.c_int32),
                ("optimized_sampling", C.c_int32), ("reserved", C.c_int32)]


_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double

# every symbol declared in include/morbit_rbf.h with its argument types
SIGNATURES = {
    "mrbf_abi_version": (C.c_int, []),
    "mrbf_init": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "mrbf_set_stream": (C.c_int, [_vp, _vp]),
    "mrbf_get_stream": (C.c_int, [_vp, C.POINTER(_vp)]),
    "mrbf_sync": (C.c_int, [_vp]),
    "mrbf_set_isapprox_rtol": (C.c_int, [_vp, _f64]),
    "mrbf_destroy": (None, [_vp]),
    "mrbf_last_error": (C.c_char_p, [_vp]),
    "mrbf_launch_count": (_i64, [_vp]),
    "mrbf_profile_enable": (C.c_int, [_vp, _i32]),
    "mrbf_profile_read": (C.c_int, [_vp, _vp]),
    "mrbf_select_points": (C.c_int, [_vp, C.POINTER(MrbfCfg), _i32, _i32, _i32] + [_vp] * 4 + [_vp, _f64] + [_vp] * 4
                           + [_vp] * 6 + [_i32] + [_vp] * 6),
    "mrbf_select_points_dev": (C.c_int, [_vp, C.POINTER(MrbfCfg), _i32, _i32, _i32] + [_vp] * 4 + [_vp, _f64] + [_vp] * 4
                               + [_vp] * 6 + [_i32] + [_vp] * 6),
    "mrbf_select_points_keep_dev": (C.c_int, [_vp, C.POINTER(MrbfCfg), _i32, _i32, _i32] + [_vp] * 4 + [_vp, _f64] + [_vp] * 4
                                    + [_vp] * 6 + [_i32] + [_vp] * 6 + [C.POINTER(_vp)]),
    "mrbf_select_points_keep": (C.c_int, [_vp, C.POINTER(MrbfCfg), _i32, _i32, _i32] + [_vp] * 4 + [_vp, _f64] + [_vp] * 4
                                + [_vp] * 6 + [_i32] + [_vp] * 6 + [C.POINTER(_vp)]),
    "mrbf_build_prepared": (C.c_int, [_vp, C.POINTER(MrbfCfg), _vp, _i32] + [_vp] * 10 + [C.POINTER(_vp), _vp]),
    "mrbf_free_prepared": (None, [_vp, _vp]),
    "mrbf_build_prepared_dev": (C.c_int, [_vp, C.POINTER(MrbfCfg), _vp, _i32] + [_vp] * 10 + [C.POINTER(_vp), _vp]),
    "mrbf_round4": (C.c_int, [_vp, C.POINTER(MrbfCfg), _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _i32, _vp, _vp,
                              _i32, _vp, _vp, _vp]),
    "mrbf_gather_training_dev": (C.c_int, [_vp, _i32, _i32, _i32, _i32] + [_vp] * 10 + [_i32, _vp, _vp, _i32, _vp, _vp, _vp]),
    "mrbf_build": (C.c_int, [_vp, C.POINTER(MrbfCfg), _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, C.POINTER(_vp), _vp]),
    "mrbf_build_dev": (C.c_int, [_vp, C.POINTER(MrbfCfg), _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, C.POINTER(_vp), _vp]),
    "mrbf_free_model": (None, [_vp, _vp]),
    "mrbf_model_dims": (C.c_int, [_vp, _vp]),
    "mrbf_model_coeffs": (C.c_int, [_vp, _vp, _vp, _vp]),
    "mrbf_eval": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "mrbf_eval_dev": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "mrbf_backtrack": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _f64, _f64, _f64, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "mrbf_backtrack_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _f64, _f64, _f64, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "mrbf_descent_direction": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "mrbf_descent_direction_dev": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "mrbf_ps_solve": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp]),
    "mrbf_ps_solve_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _vp]),
    "mrbf_db_append_dev": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "mrbf_model_scatter_dev": (C.c_int, [_vp, _vp, _vp, _vp, _i32]),
    "mrbf_comm_unique_id": (C.c_int, [_vp]),
    "mrbf_comm_init": (C.c_int, [C.c_int, _vp, _i32, _i32, C.POINTER(_vp)]),
    "mrbf_comm_from_nccl": (C.c_int, [C.c_int, _vp, _i32, _i32, C.POINTER(_vp)]),
    "mrbf_comm_destroy": (None, [_vp]),
    "mrbf_comm_last_error": (C.c_char_p, [_vp]),
    "mrbf_gather": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp]),
}

_lib = None


def load():
    """Load libmorbit_rbf.so (no compute happens here).  Raises ImportError loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (needs nvcc). There is deliberately no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library ever diverge
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
