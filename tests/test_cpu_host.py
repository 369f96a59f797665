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


def _header_prototypes():
    """{name: (ret, [kinds])} parsed from include/morbit_rbf.h; kinds: ptr / i32 / i64 / f64."""
    header = open(os.path.join(ROOT, "include", "morbit_rbf.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|void|int64_t|const char\s*\*)\s*(mrbf_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3)
        kinds = []
        for a in [a.strip() for a in args.split(",")]:
            if a in ("", "void"):
                continue
            if "*" in a:
                kinds.append("ptr")
            elif re.match(r"(const\s+)?double\b", a):
                kinds.append("f64")
            elif re.match(r"(const\s+)?int64_t\b", a):
                kinds.append("i64")
            elif re.match(r"(const\s+)?(int32_t|int)\b", a):
                kinds.append("i32")
            else:
                raise AssertionError(f"unparsed argument {a!r} of {name}")
        protos[name] = ({"int": "i32", "void": "void", "int64_t": "i64"}.get(ret, "ptr"), kinds)
    return protos


def test_ctypes_signatures_match_header_prototypes():
    """The ctypes table the tests call through (morbit.jl_b200/_lib.py) against the prototypes in include/morbit_rbf.h:
    same entry points, same number of arguments, pointer / int32 / int64 / double in the same places."""
    import ctypes as C
    import morbit_jl_b200 as mb
    protos = _header_prototypes()
    assert set(protos) == set(mb._lib.SIGNATURES), set(protos) ^ set(mb._lib.SIGNATURES)

    def kind(t):
        if t in (C.c_int32, C.c_int):
            return "i32"
        if t is C.c_int64:
            return "i64"
        if t is C.c_double:
            return "f64"
        return "ptr"
    for name, (res, args) in mb._lib.SIGNATURES.items():
        ret, kinds = protos[name]
        assert [kind(a) for a in args] == kinds, (name, [kind(a) for a in args], kinds)
        assert ("void" if res is None else kind(res)) == ret, (name, res, ret)


def test_julia_shim_ccall_signatures_match_header():
    """Every `ccall((:mrbf_..., LIBMRBF), Ret, (T...), ...)` in morbit.jl_b200/julia/GpuRbf.jl (the binding a Morbit maintainer adds;
    Julia cannot run in this image) against the header's prototypes: existing symbol, argument count, pointer / Int32 / Int64 /
    Float64 in the same places, and as many call arguments as declared types."""
    src = open(os.path.join(ROOT, "morbit.jl_b200", "julia", "GpuRbf.jl")).read()
    protos = _header_prototypes()
    calls = list(re.finditer(r"ccall\(\(:(mrbf_\w+), LIBMRBF\),\s*(\w+),\s*\(([^)]*)\)", src, flags=re.S))
    assert len(calls) >= 14
    jl_kind = lambda t: ("ptr" if t.startswith(("Ptr{", "Ref{")) or t == "Cstring" else
                         {"Int32": "i32", "Cint": "i32", "Int64": "i64", "Float64": "f64"}[t])
    seen = set()
    for m in calls:
        name, ret, tup = m.group(1), m.group(2), m.group(3)
        assert name in protos, f"{name} is not declared in morbit_rbf.h"
        types = [t.strip() for t in tup.replace("\n", " ").split(",") if t.strip()]
        assert [jl_kind(t) for t in types] == protos[name][1], (name, types, protos[name][1])
        assert {"Cint": "i32", "Cvoid": "void", "Cstring": "ptr", "Int64": "i64"}[ret] == protos[name][0], (name, ret)
        # number of values passed == number of declared types (count top-level commas up to the matching parenthesis)
        i, depth, nargs, cur = m.end(), 1, 0, ""
        while depth > 0:
            ch = src[i]
            if ch in "([{":
                depth += 1
            elif ch in ")]}":
                depth -= 1
                if depth == 0:
                    break
            if ch == "," and depth == 1:
                nargs += 1 if cur.strip() else 0
                cur = ""
            else:
                cur += ch
            i += 1
        nargs += 1 if cur.strip() else 0
        assert nargs == len(types), (name, nargs, len(types))
        seen.add(name)
    # the shim binds the fused path the benchmark measures, the descent hooks and the gather
    for need in ("mrbf_select_points_keep", "mrbf_build_prepared", "mrbf_build", "mrbf_round4", "mrbf_eval", "mrbf_backtrack",
                 "mrbf_descent_direction", "mrbf_set_isapprox_rtol", "mrbf_gather", "mrbf_comm_init", "mrbf_free_prepared"):
        assert need in seen, need
    # struct MrbfCfg mirrors struct mrbf_cfg field by field
    header = open(os.path.join(ROOT, "include", "morbit_rbf.h")).read()
    c_fields = re.findall(r"^\s*(int32_t|double)\s+(\w+);", header[header.index("typedef struct mrbf_cfg {"):header.index("} mrbf_cfg;")], flags=re.M)
    jl_fields = re.findall(r"^\s*(\w+)::(Int32|Float64)\s*$", src[src.index("struct MrbfCfg"):src.index("const KERNEL_IDS")], flags=re.M)
    assert [(n, {"int32_t": "Int32", "double": "Float64"}[t]) for t, n in c_fields] == [(n, t) for n, t in jl_fields]


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


def test_stream_ordered_lib_brackets_only_the_device_entry_points():
    """engine._StreamOrderedLib: symbols that do not end in `_dev` are the library's own; `_dev` calls are bracketed by stream waits only
    when the context's stream differs from torch's current one (exercised on the GPU by
    test_dev_entry_points_are_ordered_with_torchs_stream; here the control flow with stand-ins)."""
    from morbit_jl_b200.engine import _StreamOrderedLib

    class FakeLib:
        def __init__(self): self.calls = []
        def mrbf_sync(self, *a): self.calls.append(("mrbf_sync", a)); return 0
        def mrbf_eval_dev(self, *a): self.calls.append(("mrbf_eval_dev", a)); return 7

    class FakeStream:
        def __init__(self, log, name): self.log, self.name = log, name
        def wait_stream(self, other): self.log.append((self.name, "waits for", other.name))

    class FakeEngine:
        device = 0
        def __init__(self): self.ext = None
        def _foreign_stream(self): return self.ext

    lib, eng = FakeLib(), FakeEngine()
    proxy = _StreamOrderedLib(lib, eng)
    assert proxy.mrbf_sync.__self__ is lib and proxy.mrbf_sync(1) == 0            # passed through untouched
    assert proxy.mrbf_eval_dev(1, 2) == 7 and lib.calls[-1] == ("mrbf_eval_dev", (1, 2))     # same stream: a plain call
    # a foreign stream: engine stream waits for torch's, then torch's waits for the engine's -- needs torch only for current_stream()
    import torch
    log = []
    eng.ext = FakeStream(log, "engine")
    cur = FakeStream(log, "torch")
    orig = torch.cuda.current_stream
    torch.cuda.current_stream = lambda device=None: cur
    try:
        assert proxy.mrbf_eval_dev(3) == 7
    finally:
        torch.cuda.current_stream = orig
    assert log == [("engine", "waits for", "torch"), ("torch", "waits for", "engine")] and lib.calls[-1] == ("mrbf_eval_dev", (3,))
    with pytest.raises(AttributeError):
        proxy.mrbf_no_such_symbol
