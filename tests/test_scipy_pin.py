"""Independent pin of the oracle's model build / evaluation.

The reference's coefficient solve lives in the un-vendored RadialBasisFunctionModels.jl and ships no golden
vectors, so the oracle cannot be pinned by the reference itself.  `scipy.interpolate.RBFInterpolator` is an
implementation of the same mathematical object written by other people: the augmented saddle-point system
[Phi Pi; Pi' 0][w; lam] = [Y; 0] with monomial tail of total degree <= `degree`, solved by LAPACK gesv.
Its kernels `cubic` (r^3), `linear` (-r), `quintic` (-r^5), `multiquadric` (-sqrt(1 + (eps r)^2)),
`inverse_multiquadric`, `gaussian` (exp(-(eps r)^2)) and `thin_plate_spline` (r^2 log r) are, sign included, the
radial functions of the oracle's assumption register U1-U5.  Agreement of values to cond * eps
pins: Euclidean rho (U1), the kernel formulas and the alpha convention (U2), the tail basis / degree (U4), the
saddle system (U5) and -- through finite differences of SciPy's values -- the Jacobian formula incl. rho = 0 (U6).
"""
import numpy as np
import pytest
from scipy.interpolate import RBFInterpolator

from oracle import c_oracle as CO
from oracle import rbf_oracle as O

# (oracle kernel, oracle shape parameter, scipy kernel, scipy epsilon)
PAIRS = [
    ("cubic", float("nan"), "cubic", 1.0),
    ("cubic", 1.0, "linear", 1.0),
    ("cubic", 5.0, "quintic", 1.0),
    ("multiquadric", float("nan"), "multiquadric", 1.0),
    ("multiquadric", 2.5, "multiquadric", 2.5),
    ("inv_multiquadric", 0.7, "inverse_multiquadric", 0.7),
    ("gaussian", float("nan"), "gaussian", 1.0),
    ("gaussian", 1.7, "gaussian", 1.7),
    ("thin_plate_spline", 1.0, "thin_plate_spline", 1.0),
]


def _data(rng, n, N, k):
    S = rng.random((N, n))
    V = np.stack([np.sum(S**2, 1), np.sum(np.sin(3 * S), 1), S[:, 0] * S[:, -1]], 1)[:, :k]
    return S, V


@pytest.mark.parametrize("kernel,shape,sk,eps", PAIRS)
@pytest.mark.parametrize("n,N,k", [(2, 12, 2), (5, 30, 3), (10, 66, 1), (30, 61, 2)])
def test_oracle_build_matches_scipy_rbfinterpolator(kernel, shape, sk, eps, n, N, k):
    rng = np.random.default_rng(1000 * n + N)
    S, V = _data(rng, n, N, k)
    cfg = O.RbfConfig(kernel=kernel, shape_parameter=shape, polynomial_degree=1)
    if O.get_radial_function(cfg).cpd_order > 2:
        pytest.skip("quintic needs a quadratic tail (degree > 1 is outside the hot path)")
    m = O.build_model(S, V, cfg)
    ref = RBFInterpolator(S, V, kernel=sk, epsilon=eps, degree=1, smoothing=0.0)
    X = np.vstack((rng.random((25, n)), S[:3], S[:2] + 1e-3))
    Yr = ref(X)
    Yo = np.array([m.eval(x) for x in X])
    tol = max(1e-10, 50 * m.cond * np.finfo(float).eps)      # two LU solves of the same system: O(cond * eps)
    scale = max(1.0, np.abs(Yr).max())
    assert np.abs(Yo - Yr).max() <= tol * scale, (np.abs(Yo - Yr).max() / scale, m.cond)
    # the C twin evaluates the same model
    w, lam, st = CO.build_batched(cfg, S[None], V[None], [N])
    assert st[0] == 0
    Yc = CO.eval_points(cfg, S, w[0], lam[0], X)
    assert np.abs(Yc - Yr).max() <= tol * scale, (np.abs(Yc - Yr).max() / scale, m.cond)
    # Jacobian of the oracle against central differences of SciPy's values (SciPy offers no derivative)
    h = 1e-5
    Xd = X[:6]
    Jo = np.array([m.jac(x) for x in Xd])
    Jfd = np.zeros_like(Jo)
    for c in range(n):
        e = np.zeros(n); e[c] = h
        Jfd[:, :, c] = (ref(Xd + e) - ref(Xd - e)) / (2 * h)
    jscale = max(1.0, np.abs(Jo).max())
    fd_tol = 2e-6 if kernel != "cubic" or shape != 1.0 else 2e-5        # -r has a kink at the sites
    assert np.abs(Jo - Jfd).max() <= max(fd_tol, 1e3 * tol) * jscale, np.abs(Jo - Jfd).max() / jscale


@pytest.mark.parametrize("kernel,shape,sk,eps", [("gaussian", 1.3, "gaussian", 1.3), ("inv_multiquadric", float("nan"), "inverse_multiquadric", 1.0)])
@pytest.mark.parametrize("deg", [-1, 0])
def test_oracle_low_degree_tails_match_scipy(kernel, shape, sk, eps, deg):
    """Positive definite kernels with no tail / a constant tail (polynomial_degree -1 and 0, test/rbf_models.jl:29)."""
    rng = np.random.default_rng(5 + deg)
    S, V = _data(rng, 4, 25, 2)
    cfg = O.RbfConfig(kernel=kernel, shape_parameter=shape, polynomial_degree=deg)
    m = O.build_model(S, V, cfg)
    ref = RBFInterpolator(S, V, kernel=sk, epsilon=eps, degree=deg)
    X = rng.random((20, 4))
    Yo = np.array([m.eval(x) for x in X])
    tol = max(1e-10, 50 * m.cond * np.finfo(float).eps)
    assert np.abs(Yo - ref(X)).max() <= tol * max(1.0, np.abs(Yo).max())


def test_degree_is_raised_to_cpd_minus_one_like_scipy_requires():
    """U4: a cubic model asked for degree -1 / 0 is built with the linear tail (SciPy refuses lower degrees only with a warning;
    its minimum degree table is the same cpd_order - 1)."""
    rng = np.random.default_rng(9)
    S, V = _data(rng, 3, 15, 1)
    m = O.build_model(S, V, O.RbfConfig(kernel="cubic", polynomial_degree=-1))
    assert m.degree == 1
    ref = RBFInterpolator(S, V, kernel="cubic", degree=1)
    X = rng.random((10, 3))
    assert np.abs(np.array([m.eval(x) for x in X]) - ref(X)).max() <= 1e-9


@pytest.mark.parametrize("n,N", [(2, 15), (4, 40)])
def test_quadratic_tail_matches_scipy_quintic(n, N):
    """cubic with exponent 5 is SciPy's `quintic` (-rho^5), conditionally positive definite of order 3: the oracle raises the tail to
    degree 2 (U4) -- the same saddle system as RBFInterpolator(kernel="quintic", degree=2).  Pins the quadratic tail and the sign."""
    from scipy.interpolate import RBFInterpolator
    rng = np.random.default_rng(n + N)
    S = rng.random((N, n)); V = np.stack([np.sum(S ** 2, -1), np.sum(np.sin(3 * S), -1)], -1); X = rng.random((9, n))
    m = O.build_model(S, V, O.RbfConfig(kernel="cubic", shape_parameter=5.0))
    assert m.degree == 2 and m.lam.shape[0] == (n + 1) * (n + 2) // 2
    Yo = np.array([m.eval(x) for x in X])
    Ys = RBFInterpolator(S, V, kernel="quintic", degree=2)(X)
    assert np.abs(Yo - Ys).max() <= 1e-9 * np.abs(Ys).max()
    h = 1e-6
    Jfd = np.array([(m.eval(X[0] + h * np.eye(n)[i]) - m.eval(X[0] - h * np.eye(n)[i])) / (2 * h) for i in range(n)]).T
    assert np.abs(m.jac(X[0]) - Jfd).max() <= 1e-6 * max(1.0, np.abs(Jfd).max())
