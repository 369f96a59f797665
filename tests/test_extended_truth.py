"""oracle/extended.py (extended-precision solution of the saddle system) against the float64 oracle on well-conditioned systems."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import rbf_oracle as O
from oracle.extended import truth_values


@pytest.mark.parametrize("kernel", ["cubic", "multiquadric"])
def test_truth_agrees_with_float64_oracle_on_a_well_conditioned_system(kernel):
    rng = np.random.default_rng(5)
    n, N = 6, 40
    S = rng.random((N, n)); V = np.stack([np.sum(S ** 2, -1), np.sum(np.sin(3 * S), -1)], -1); X = rng.random((7, n))
    Y, J, info = truth_values(kernel, 1.0, S, V, X)
    om = O.build_model(S, V, O.RbfConfig(kernel=kernel))
    Yo = np.array([om.eval(x) for x in X]); Jo = np.array([om.jac(x) for x in X])
    assert info["residual"] < 1e-15 and info["cond"] < 1e6
    assert float(np.max(np.abs(Yo - Y))) <= 1e-12 * float(np.max(np.abs(Y)))
    assert float(np.max(np.abs(Jo - J))) <= 1e-11 * float(np.max(np.abs(J)))
    # interpolation: the extended-precision model reproduces the data far below float64 rounding of the values
    Yt, _, _ = truth_values(kernel, 1.0, S, V, S[:5])
    assert float(np.max(np.abs(Yt - V[:5]))) <= 1e-15 * info["cond"]
