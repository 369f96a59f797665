"""CPU-side checks: the C-ABI library loads and exports every symbol include/morbit_rbf.h declares (no compute
without a GPU), host bookkeeping, sharding, and the world_size-2 gloo gather."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    import morbit_jl_b200 as mb
    header = open(os.path.join(ROOT, "include", "morbit_rbf.h")).read()
    declared = set(re.findall(r"\b(mrbf_[a-z0-9_]+)\s*\(", header))
    declared -= {"mrbf_cfg", "mrbf_ctx", "mrbf_model"}
    lib = ctypes.CDLL(mb.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in morbit_rbf.h but not exported"
    assert declared == set(mb._lib.SIGNATURES), declared ^ set(mb._lib.SIGNATURES)
    assert lib.mrbf_abi_version() == 2


def test_no_cpu_fallback_without_gpu():
    import torch
    import morbit_jl_b200 as mb
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mb.MrbfError):
        mb.Engine(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "morbit.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "rbf_oracle" not in txt and "c_oracle" not in txt and "from oracle" not in txt, f


def test_config_asserts_and_defaults():
    import morbit_jl_b200 as mb
    cfg = mb.RbfConfig()
    assert cfg.kernel == "cubic" and cfg.theta_pivot == 0.25 and cfg.polynomial_degree == 1
    assert cfg.signature() == (0.25, 2.0, 2.0, True)
    assert mb.combinable(cfg) and mb.max_evals(cfg) == 2**63 - 1
    assert mb.RbfConfig(kernel="cubic") == mb.RbfConfig(kernel="cubic") and hash(mb.RbfConfig()) == hash(mb.RbfConfig())
    for bad in (dict(kernel="exp"), dict(kernel="cubic", shape_parameter=2.0), dict(theta_enlarge_1=0.5),
                dict(theta_pivot=0.9), dict(kernel="gaussian", shape_parameter=-1.0)):
        with pytest.raises(AssertionError):
            mb.RbfConfig(**bad)
    assert mb.max_model_points(cfg, 30) == 496 and mb.max_model_points(mb.RbfConfig(max_model_points=61), 30) == 61
    c = mb.to_c_cfg(mb.RbfConfig(kernel="gaussian", shape_parameter=2.0))
    assert c.kernel == 4 and c.shape_parameter == 2.0
    assert mb.surrogate.parse_shape_param_string(0.5, "10/Δ") == 20.0


def test_arraydb_bookkeeping():
    import morbit_jl_b200 as mb
    db = mb.ArrayDB(3)
    a = db.new_result([0.1, 0.2, 0.3], [1.0])
    b = db.new_result([0.4, 0.5, 0.6], None)
    c = db.new_result([0.7, 0.8, 0.9], [])
    assert (a, b, c) == (1, 2, 3) and db.unevaluated_ids == [2, 3] and db.num_entries == 3
    assert db.find_result([0.4, 0.5, 0.6]) == 2 and db.find_result([9, 9, 9]) == -1
    assert db.ensure_contains_res_with_site([0.7, 0.8, 0.9]) == 3
    assert db.ensure_contains_res_with_site([1.0, 1.0, 1.0]) == 4 and db.unevaluated_ids == [2, 3, 4]
    for i in range(100):
        db.new_result(np.full(3, i), [float(i)])
    assert db.num_entries == 104 and db.sites_array().shape == (104, 3)
    assert db.eval_missing(lambda x: [x.sum()]) == 3 and db.unevaluated_ids == []
    np.testing.assert_allclose(db.get_value(2), [1.5])


def test_shard_range_partitions():
    from morbit_jl_b200.multistart import shard_range
    for total in (1, 7, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_synthetic_inputs_are_seeded():
    from morbit_jl_b200 import synthetic
    a = synthetic.multistart_batch(3, n=5, n_db=8)
    b = synthetic.multistart_batch(2, n=5, n_db=8, first_instance=1)
    np.testing.assert_array_equal(a["sites"][1:], b["sites"])
    assert np.all(a["sites"] >= 0) and np.all(a["sites"] <= 1) and np.array_equal(a["sites"][:, 0], a["x"])
    h = synthetic.halton(4, 3)
    np.testing.assert_allclose(h[:3, 0], [0.5, 0.25, 0.75]); np.testing.assert_allclose(h[0, 1], 1 / 3)


_GLOO_SCRIPT = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch.distributed as dist
from morbit_jl_b200.multistart import shard_range, gather_results
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
total = 11
lo, hi = shard_range(total, rank, world)
local = np.arange(lo, hi, dtype=np.float64)[:, None] * np.array([1.0, 10.0, 100.0])[None, :]
out = gather_results(local, total, rank, world)
exp = np.arange(total, dtype=np.float64)[:, None] * np.array([1.0, 10.0, 100.0])[None, :]
assert out.shape == (total, 3) and np.array_equal(out, exp), out
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_gather_results_gloo_world2(tmp_path):
    script = tmp_path / "gloo_gather.py"
    script.write_text(_GLOO_SCRIPT)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out
