import sys, math, numpy as np
sys.path.insert(0, "/root/repo")
import morbit_jl_b200 as mb
from morbit_jl_b200 import lockstep as L, synthetic
from oracle import rbf_oracle as O, iterate_oracle as IO, descent_oracle as D
rng = np.random.default_rng(1)
B, n = 12, 2
x0 = np.vstack([[-np.pi, 2.71828], rng.uniform(-3, 3, (B - 1, n))])
glb, gub = np.full(n, -np.inf), np.full(n, np.inf)
MAXIT = 25
if len(sys.argv) > 1 and sys.argv[1] == "zdt3":
    B, n, MAXIT = 6, 30, 6
    x0 = synthetic.halton(B, n); glb, gub = np.zeros(n), np.ones(n)
    FUNC = lambda z: synthetic.zdt3(np.asarray(z)); OCFG = O.RbfConfig(kernel="multiquadric")
    drv = L.LockstepDriver(mb.RbfConfig(kernel="multiquadric"), synthetic.zdt3, x0, glb, gub, L.AlgorithmConfig(max_iter=MAXIT), capacity=128, record=True)
else:
    FUNC = lambda z: synthetic.two_parabolas(np.asarray(z)); OCFG = O.RbfConfig(kernel="cubic")
    drv = L.LockstepDriver(mb.RbfConfig(kernel="cubic"), synthetic.two_parabolas, x0, glb, gub, L.AlgorithmConfig(max_iter=MAXIT), capacity=128, record=True)
drv.run()
S = drv.sites.cpu().numpy(); V = drv.values.cpu().numpy()
def replay(b, verbose=False):
    calls = [(c["d"][b], float(c["omega"][b])) for c in drv.lp_calls if c["mask"][b]]
    ci = [0]
    log = []
    def feed(run, J):
        d, om = calls[ci[0]]; ci[0] += 1
        dh, oh = D.lp_highs(run.x, J, glb, gub, True)
        log.append("   lp call %d gpu omega %r highs omega %r gpu d %s highs d %s" % (ci[0], om, oh, d, dh))
        return d, om
    run = IO.Run(FUNC, x0[b], glb, gub, OCFG, IO.AlgoConfig(max_iter=MAXIT), direction=feed)
    for t, tr in enumerate(drv.trace):
        if run.ret_code != 0:
            break
        err = None
        try:
            r = run.iterate()
        except (AssertionError, IndexError) as e:
            err = repr(e)
            r = run.records[-1] if run.records else None
        nd = min(run.db.num_entries, S.shape[1])
        ds = np.abs(np.array(run.db.sites)[:nd] - S[b, :nd]).max(1)
        line = (t, "gpu", tr["ret"][b], tr["it_stat"][b], tr["x_index"][b], tr["n_db"][b], tr["delta"][b], tr["omega"][b], tr["rho"][b], bool(tr["fully_linear"][b]),
              "| or", run.ret_code, run.it_stat, run.x_index, run.db.num_entries, run.delta, r.omega if r else None, r.rho if r else None, run.meta.fully_linear, r.n_crit_loops if r else None, "| max site diff", ds.max(), "first bad row", int(np.argmax(ds > 1e-9)) if (ds > 1e-9).any() else -1)
        log.append(" ".join(str(v) for v in line))
        bad = err or (ds > 1e-9).any() or tr["ret"][b] != run.ret_code or tr["it_stat"][b] != run.it_stat or tr["n_db"][b] != run.db.num_entries
        if bad:
            print("instance", b, "DIVERGES at iteration", t, err)
            print("\n".join(log[-8:]))
            return False
    print("instance", b, "ok", len(run.records), "iterations, ret", run.ret_code)
    return True
for b in range(B):
    replay(b)

def select_debug(b, t, efl=False):
    """state of instance b before iteration t (0-based) from the GPU trace; compare select of GPU / numpy oracle / C oracle."""
    from oracle import c_oracle as CO
    tr = drv.trace[t - 1]
    nd = int(tr["n_db"][b]); x = tr["x"][b]; xi = int(tr["x_index"][b]); delta = float(tr["delta"][b])
    sites = S[b, :nd]
    cfg = mb.RbfConfig(kernel="cubic") if OCFG.kernel == "cubic" else mb.RbfConfig(kernel=OCFG.kernel)
    eng = drv.engine
    res = eng.select_points(cfg, sites[None], [nd], [xi], x[None], [delta], 0.5, glb, gub, efl, False)
    print("GPU  r1", res.r1[0, :res.n_r1[0]], "r2", res.r2[0, :res.n_r2[0]], "n_r3", res.n_r3[0], "r3", res.r3_sites[0, :res.n_r3[0]], "r4", res.r4[0, :res.n_r4[0]], "fl", res.flags_out[0], "dirs", res.dirs[0, :res.n_dirs[0]])
    ref = CO.select_points_batched(cfg, sites[None], [xi], x[None], [delta], 0.5, glb, gub, efl, False, 2**31 - 1)
    print("C    r1", ref.r1[0, :ref.n_r1[0]], "r2", ref.r2[0, :ref.n_r2[0]], "n_r3", ref.n_r3[0], "r3", ref.r3_sites[0, :ref.n_r3[0]], "r4", ref.r4[0, :ref.n_r4[0]], "fl", ref.fully_linear[0], "dirs", ref.dirs[0, :ref.n_dirs[0]])
    db = O.ArrayDB()
    for s_ in sites: db.new_result(s_, np.zeros(2))
    meta = O.RbfMeta(signature=OCFG.signature())
    tf = O.FilterTrace() if hasattr(O, "FilterTrace") else None
    meta = O.prepare_update_model(meta, OCFG, db, x, xi, delta, 0.5, glb, gub, ensure_fully_linear=efl, num_objf_evals=nd, trace=tf)
    print("NP   r1", meta.round1_indices, "r2", meta.round2_indices, "r3", [db.get_site(i) for i in meta.round3_indices], "r4", meta.round4_indices, "fl", meta.fully_linear, "dirs", meta.improving_directions)
    print("x", x, "delta", delta, "nd", nd)
    if tf is not None: print("filter trace", tf.__dict__)
if len(sys.argv) > 3:
    np.set_printoptions(linewidth=250, precision=4)
    select_debug(int(sys.argv[2]), int(sys.argv[3]), len(sys.argv) > 4)
