"""CPU checks of the scalar `iterate!` restatement (oracle/iterate_oracle.py) that the lock-step GPU driver is compared with:
properties the reference's algorithm guarantees (no golden vectors exist for it: parity unpinned)."""
import numpy as np

from oracle import rbf_oracle as O, iterate_oracle as IO


def _two_parabolas(x):
    x = np.asarray(x)
    return np.array([np.sum((x - 1.0) ** 2), np.sum((x + 1.0) ** 2)])


def test_float32_defaults_of_the_reference():
    """AbstractConfigInterface.jl:14-95: the defaults are Float32 literals promoted to Float64."""
    ac = IO.AlgoConfig()
    assert ac.delta_0 == float(np.float32(0.1)) != 0.1 and ac.delta_max == 0.5 and ac.gamma_shrink_much == float(np.float32(0.51))
    assert abs(ac.f_tol_rel - 3.4526698e-4) < 1e-10 and abs(ac.omega_tol_rel - 3.4526698e-3) < 1e-9 and ac.mu == 2000.0


def test_two_parabolas_reaches_the_pareto_set():
    """examples/example_two_parabolas.jl: from (-pi, e) the run ends on the Pareto set {x1 = x2 in [-1, 1]}."""
    r = IO.optimize(_two_parabolas, np.array([-np.pi, 2.71828]), np.full(2, -np.inf), np.full(2, np.inf), O.RbfConfig(kernel="cubic"),
                    IO.AlgoConfig(max_iter=40))
    assert r.ret_code in (IO.TOLERANCE, IO.CRITICAL, IO.MAX_ITER)
    assert abs(r.x[0] - r.x[1]) < 0.02 and np.all(np.abs(r.x) <= 1.05)
    assert r.db.num_entries == r.num_evals                              # every evaluation is a database row
    fx = np.array([rec.fx for rec in r.records])
    assert np.all(np.diff(fx, axis=0) <= 1e-12)                         # strict acceptance, nu_accept = 0: no objective ever increases
    deltas = [rec.delta for rec in r.records]
    assert max(deltas) <= 0.5 and deltas[0] == 2 * float(np.float32(0.1))   # first step successful: radius doubled (gamma_grow)
    # a rejected step with a fully linear model shrinks by gamma_shrink_much, the iterate stays
    for a, b in zip(r.records[:-1], r.records[1:]):
        if b.it_stat == IO.INACCEPTABLE:
            assert b.x_index == a.x_index and abs(b.delta - a.delta * float(np.float32(0.51))) < 1e-15
        if b.it_stat == IO.MODELIMPROVING:
            assert b.x_index == a.x_index and b.delta == a.delta


def test_stop_codes():
    f, x0, inf = _two_parabolas, np.array([2.0, -1.5]), np.full(2, np.inf)
    r = IO.optimize(f, x0, -inf, inf, O.RbfConfig(kernel="cubic"), IO.AlgoConfig(max_iter=3))
    assert r.ret_code == IO.MAX_ITER and len(r.records) == 4 and r.records[-1].it_stat == IO.EARLY_EXIT
    r = IO.optimize(f, x0, -inf, inf, O.RbfConfig(kernel="cubic"), IO.AlgoConfig(max_evals=6))
    assert r.ret_code == IO.BUDGET_EXHAUSTED and r.num_evals <= 7
    r = IO.optimize(f, np.array([0.3, 0.3]), -inf, inf, O.RbfConfig(kernel="cubic"), IO.AlgoConfig(max_iter=60))
    assert r.ret_code in (IO.CRITICAL, IO.TOLERANCE)                   # started on the Pareto set


def test_boxed_run_stays_feasible():
    glb, gub = np.array([-0.5, 0.2]), np.array([0.4, 3.0])
    r = IO.optimize(_two_parabolas, np.array([5.0, 2.5]), glb, gub, O.RbfConfig(kernel="multiquadric"), IO.AlgoConfig(max_iter=25))
    assert np.all(r.x >= glb) and np.all(r.x <= gub) and r.records[0].x_index >= 1
    for s in r.db.sites:
        assert np.all(s >= glb - 1e-15) and np.all(s <= gub + 1e-15)
