"""Print the key figures of a bench.py JSON line: python tools/bench_brief.py gpurun_out/bench.json"""
import json, sys
for line in open(sys.argv[1]):
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    if "value" not in d:
        print(d); continue
    print(f"{d.get('impl', 'gpu')}: {d['value']:.0f} {d['unit']}  {d['ms_per_step']:.3f} ms/step  n_gpus={d['n_gpus']}")
    e = d.get("e2e") or {}
    print("  e2e", round(e.get("value", 0)), "h2d", e.get("h2d_bytes_per_step"), "snapshot", round((d.get("e2e_snapshot") or {}).get("value", 0)),
          "resident", round((d.get("e2e_resident") or {}).get("value", 0)))
    r = d.get("roofline") or {}
    print("  roofline", r.get("kernel"), round(r.get("frac", 0), 4), "kernel_ms", r.get("kernel_ms"))
    c = d.get("cpu_baseline") or {}
    print("  cpu", round(c.get("value", 0)), c.get("cores"))
    for s in d.get("secondary", []):
        rf = (s.get("roofline") or {}).get("frac")
        print("   -", s["metric"], s.get("kernel", s.get("variant", "")), round(s["value"], 1), s.get("unit"), "frac" if rf else "", round(rf, 3) if rf else "")
