"""CPU: the two restatements of the steepest-descent LP (descent.jl:75-135) agree with each other."""
import numpy as np

from oracle import descent_oracle as D


def test_k2_breakpoint_enumeration_matches_highs():
    rng = np.random.default_rng(0)
    for trial in range(40):
        n = int(rng.integers(2, 40))
        jac = rng.normal(size=(2, n)) * rng.choice([1.0, 1e-3, 50.0])
        lb, ub = np.zeros(n), np.ones(n)
        x = rng.random(n)
        if trial % 3 == 0:
            x[rng.integers(0, n)] = 0.0           # on the boundary: a one-sided box in that coordinate
        d, om = D.lp_highs(x, jac, lb, ub)
        om2 = D.lp_k2_exact(x, jac, lb, ub)
        assert abs(om - om2) <= 1e-9 * max(1.0, abs(om)), (trial, om, om2)
        D.check_optimal(x, jac, lb, ub, d, om)


def test_unconstrained_single_output_is_normalised_sign_vector():
    # k = 1, no bounds: d = -sign(g), omega = ||g||_1 / ||g||_2
    g = np.array([[3.0, -4.0, 0.5]])
    d, om = D.lp_highs(np.zeros(3), g, np.full(3, -np.inf), np.full(3, np.inf))
    assert np.allclose(d, -np.sign(g[0])) and abs(om - np.abs(g).sum() / np.linalg.norm(g)) < 1e-12
