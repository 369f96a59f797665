"""The algorithmic claim behind the hand-over of under-poised instances (Round4Params::hyb), checked on the CPU oracle alone.

Round 4 (RbfModel.jl:352-499) started from a found set of N0 < p points applies its rank guard (:433-438) only while N < p, but it
appends a column to Z on EVERY acceptance (:464-467) -- so when the set becomes poised (N = p) the reference's Z holds p - N0 directions
that are orthogonal to the constants (the first column of Q spans them) and to nothing else.  For kernels that are conditionally positive
definite of order <= 1 (Gaussian, inverse multiquadric, multiquadric) the reduced kernel matrix is positive definite on any such Z:
every candidate that is not a duplicate is accepted, exactly as by a walk that starts afresh from the poised set -- the CUDA path stops
the literal kernel there and lets the register kernels continue.  For order 2 (cubic, thin plate spline) the reference goes on rejecting
what a fresh walk accepts: no hand-over for those (mrbf_api.cu, run_round4)."""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import rbf_oracle as O


def _instance(rng, n, n_db):
    x = 0.3 + 0.4 * rng.random(n)
    lb2, ub2 = x - 0.25, x + 0.25
    sites = np.vstack([x[None], lb2 + (ub2 - lb2) * rng.random((n_db - 1, n))])
    sites[rng.choice(np.arange(1, n_db), size=max(1, n_db // 8), replace=False)] += 1.0       # some sites outside box 2
    return sites, lb2, ub2


def _walks(kernel, sites, lb2, ub2, found0, p):
    """(full walk, prefix walk capped at p points, fresh walk from the poised set or None when the prefix never gets there)"""
    n_db = len(sites)
    cfg = O.RbfConfig(kernel=kernel)
    full = [int(v) for v in CO.round4(cfg, sites, lb2, ub2, found0)[0]]
    prefix = [int(v) for v in CO.round4(O.RbfConfig(kernel=kernel, max_model_points=p), sites, lb2, ub2, found0)[0]]   # stops at N = p
    if len(found0) + len(prefix) < p:
        return full, prefix, None
    tried = np.arange(n_db) < prefix[-1]                       # 0-based indices up to the last acceptance
    keep = np.zeros(n_db, bool); keep[[f - 1 for f in found0 + prefix]] = True
    sites2 = sites.copy()
    sites2[tried & ~keep] = ub2 + 1.0                          # tried and rejected: not a candidate any more
    rest = [int(v) for v in CO.round4(cfg, sites2, lb2, ub2, found0 + prefix)[0]]
    return full, prefix, rest


def _found0(rng, sites, lb2, ub2, n_found):
    n_db = len(sites)
    f = [1] + [int(i) for i in 2 + rng.choice(n_db - 1, size=n_found - 1, replace=False)]
    return [i for i in f if np.all(lb2 <= sites[i - 1]) and np.all(sites[i - 1] <= ub2)] or [1]


@pytest.mark.parametrize("kernel", ["gaussian", "inv_multiquadric", "multiquadric"])
@pytest.mark.parametrize("n,n_db,n_found", [(2, 25, 1), (3, 40, 2), (5, 60, 1), (8, 90, 3), (12, 120, 1)])
def test_full_walk_equals_prefix_walk_plus_fresh_walk_for_cpd_order_up_to_one(kernel, n, n_db, n_found):
    rng = np.random.default_rng(100 * n + n_db + n_found)
    handed = 0
    for rep in range(4):
        sites, lb2, ub2 = _instance(rng, n, n_db)
        found0 = _found0(rng, sites, lb2, ub2, n_found)
        full, prefix, rest = _walks(kernel, sites, lb2, ub2, found0, n + 1)
        assert full[:len(prefix)] == prefix
        if rest is None:                           # never poised: the prefix walk IS the whole walk, nothing is handed over
            assert full == prefix
            continue
        handed += 1
        assert full == prefix + rest, (rep, found0)
    assert handed > 0


def test_cubic_is_the_counterexample():
    """Order 2: after the under-poised phase the reference's Z makes tau^2 negative for candidates a fresh walk accepts."""
    n, n_db, n_found = 3, 40, 2
    rng = np.random.default_rng(100 * n + n_db + n_found)
    sites, lb2, ub2 = _instance(rng, n, n_db)
    found0 = _found0(rng, sites, lb2, ub2, n_found)
    full, prefix, rest = _walks("cubic", sites, lb2, ub2, found0, n + 1)
    assert rest is not None and full[:len(prefix)] == prefix
    assert len(prefix + rest) > len(full)          # the fresh walk accepts more than the reference does
    # and the literal NumPy restatement (the reference's own dense operation order) says the same as its C twin, with tau^2 far from noise
    db = O.ArrayDB()
    for s in sites:
        db.new_result(s, [0.0])
    tr = O.Round4Trace()
    lit = [int(v) for v in O.rbf_round4(db, lb2, ub2, sites[0], 0.1, found0, O.RbfConfig(kernel="cubic"), trace=tr)]
    assert lit == full
    late = [t for t, N in zip(tr.tau2, tr.n_points) if N == n + 1]
    assert late and max(late) < -1e-4


def test_handover_property_with_the_literal_numpy_restatement():
    """Same statement with oracle/rbf_oracle.py::rbf_round4 on one small multiquadric instance."""
    rng = np.random.default_rng(7)
    n, n_db = 3, 30
    sites, lb2, ub2 = _instance(rng, n, n_db)
    x = sites[0]

    def walk(S, found, cap):
        db = O.ArrayDB()
        for s in S:
            db.new_result(s, [0.0])
        return [int(v) for v in O.rbf_round4(db, lb2, ub2, x, 0.1, found, O.RbfConfig(kernel="multiquadric", max_model_points=cap))]

    full = walk(sites, [1], -1)
    prefix = walk(sites, [1], n + 1)
    assert len(prefix) == n and full[:n] == prefix
    tried = np.arange(n_db) < prefix[-1]
    keep = np.zeros(n_db, bool); keep[[0] + [i - 1 for i in prefix]] = True
    sites2 = sites.copy(); sites2[tried & ~keep] = ub2 + 1.0
    assert full == prefix + walk(sites2, [1] + prefix, -1)
