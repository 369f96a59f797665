"""Who is closer to the exact solution of the saddle system -- the GPU build (reduced-system Cholesky / null-space QR) or the
reference's dense LU (oracle/rbf_oracle.py::build_model)?  Both against an extended-precision solution (oracle/extended.py).
    python tools/accuracy_report.py > profiles/accuracy_r02.json      (on a B200)"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CASES = [("multiquadric", 30, 61, 0.30), ("multiquadric", 30, 128, 0.30), ("multiquadric", 30, 128, 0.05), ("multiquadric", 30, 159, 0.02),
         ("cubic", 30, 128, 0.30), ("cubic", 10, 40, 2e-3), ("cubic", 10, 120, 2e-3), ("multiquadric", 10, 130, 0.01)]


def run_case(eng, kernel, n, N, box, seed=0):
    import morbit_jl_b200 as mb
    from oracle import rbf_oracle as O
    from oracle.extended import truth_values
    rng = np.random.default_rng(seed + N)
    c = 0.3 + 0.4 * rng.random(n)
    S = c + box * (rng.random((N, n)) - 0.5); S[0] = c
    V = np.stack([np.sum(S ** 2, -1), np.sum(np.sin(3 * S), -1)], -1)
    X = c + box * (rng.random((64, n)) - 0.5)
    cfg = mb.RbfConfig(kernel=kernel)
    Yt, Jt, info = truth_values(kernel, 1.0, S, V, X)
    om = O.build_model(S, V, O.RbfConfig(kernel=kernel))
    Yo = np.array([om.eval(x) for x in X]); Jo = np.array([om.jac(x) for x in X])
    rows = {}
    for route, env in (("reduced", "0"), ("qr", "1")):
        os.environ["MRBF_BUILD_GENERAL"] = env
        model, status = eng.build(cfg, S[None], V[None], [N])
        Y, J = eng.eval(model, X[None], True, True)
        model.free()
        rows[route] = (Y[0], J[0])
    os.environ["MRBF_BUILD_GENERAL"] = "0"
    sy, sj = float(np.max(np.abs(Yt))), float(np.max(np.abs(Jt)))
    err = lambda A, T, s: float(np.max(np.abs(A.astype(np.longdouble) - T)) / s)
    return dict(kernel=kernel, n=n, N=N, box=box, cond=info["cond"], truth_residual=info["residual"], refinement_steps=info["steps"],
                values=dict(lu_oracle=err(Yo, Yt, sy), gpu_reduced=err(rows["reduced"][0], Yt, sy), gpu_qr=err(rows["qr"][0], Yt, sy)),
                jacobians=dict(lu_oracle=err(Jo, Jt, sj), gpu_reduced=err(rows["reduced"][1], Jt, sj), gpu_qr=err(rows["qr"][1], Jt, sj)))


def main():
    import morbit_jl_b200 as mb
    eng = mb.Engine(0)
    out = [run_case(eng, *c) for c in CASES]
    print(json.dumps({"what": "max relative error (to the largest magnitude) of model values / Jacobians at 64 trial points against the "
                              "extended-precision solution of the saddle system; lu_oracle = the reference's dense LU in float64", "rows": out}, indent=1))


if __name__ == "__main__":
    main()
