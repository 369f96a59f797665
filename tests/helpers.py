"""Seeded corpora shared by the CPU and GPU parity tests."""
import numpy as np


def random_instances(rng, B, n, n_db, boxed=True, spread=0.6, on_bound=False):
    """B database snapshots: iterate first (id 1), the rest scattered around it at mixed distances."""
    glb = np.full(n, 0.0 if boxed else -np.inf)
    gub = np.full(n, 1.0 if boxed else np.inf)
    x = rng.random((B, n))
    if on_bound and boxed:
        x[:, 0] = 0.0
        x[:, -1] = 1.0
    sites = np.zeros((B, n_db, n))
    sites[:, 0] = x
    for b in range(B):
        r = spread * rng.random((n_db - 1, 1))
        sites[b, 1:] = np.clip(x[b] + (rng.random((n_db - 1, n)) * 2 - 1) * r, glb, gub)
    return sites, x, glb, gub


def assert_select_equal(res, ref, B, check_dirs=True):
    """res: product SelectResult (ids/counters NumPy), ref: oracle SelectResult."""
    for b in range(B):
        for name, cnt in (("r1", "n_r1"), ("r2", "n_r2"), ("r4", "n_r4")):
            a = list(getattr(res, name)[b, : getattr(res, cnt)[b]])
            e = list(getattr(ref, name)[b, : getattr(ref, cnt)[b]])
            assert a == e, f"instance {b}: {name} differs\n got {a}\n exp {e}\n margins {ref.margins[b]}"
        assert res.n_r3[b] == ref.n_r3[b], f"instance {b}: n_r3"
        assert bool(res.flags_out[b, 0]) == bool(ref.fully_linear[b]), f"instance {b}: fully_linear"
        assert bool(res.flags_out[b, 1]) == bool(ref.rebuilt[b]), f"instance {b}: rebuilt"
        k3 = ref.n_r3[b]
        if k3:
            np.testing.assert_allclose(res.r3_sites[b, :k3], ref.r3_sites[b, :k3], rtol=0, atol=1e-13)
        if check_dirs:
            assert res.n_dirs[b] == ref.n_dirs[b]
            kd = ref.n_dirs[b]
            if kd:
                np.testing.assert_allclose(res.dirs[b, :kd], ref.dirs[b, :kd], rtol=0, atol=1e-11)
