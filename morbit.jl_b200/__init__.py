"""morbit.jl_b200 -- B200-native (sm_100a) RBF-surrogate hot path of Morbit.jl behind its plugin interface.

The directory name follows the project layout (`morbit.jl_b200/`); because of the dot it is imported
through the alias module `morbit_jl_b200` at the repository root.

Importing this package loads libmorbit_rbf.so (hand-written CUDA + C ABI, include/morbit_rbf.h).  There is
no CPU fallback: a missing library raises ImportError, a missing GPU raises MrbfError at the first call.
"""
from . import _lib
from ._lib import MrbfError, LIB_PATH

_lib.load()

from .engine import Engine, Comm, ModelBatch, Prepared, SelectResult, max_model_points, to_c_cfg  # noqa: E402
from .surrogate import (  # noqa: E402
    RBF_KERNELS, RbfConfig, RbfMeta, RbfModel, ArrayDB, SuperDB, IterData, VarScaler, AlgoConfig, MopStub,
    max_evals, combinable, get_saveable, fully_linear, set_fully_linear, num_outputs, get_sub_db,
    prepare_init_model, prepare_update_model, prepare_improve_model, init_model, update_model, improve_model,
    eval_models, get_gradient, get_jacobian, _rbf_round4, _collect_indices, _backtrack, _steepest_descent_direction, default_engine,
)
from . import multistart  # noqa: E402

__all__ = [n for n in dir() if not n.startswith("__")]
