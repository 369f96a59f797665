"""Attribute ncu stall samples / executed instructions of one kernel to CUDA source lines.
The ncu SASS page has no line column, so the instruction order is zipped with `nvdisasm -g` of the same cubin.
    python tools/ncu_lines.py REPORT.ncu-rep CUBIN MANGLED_SUBSTRING [top]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, cubin, key = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, active = [], None, False
for ln in dis:
    m = re.match(r"\s*\.text\.(\S+):", ln) or re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
    if m:
        active = key in m.group(1)
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S", ln)
    if m:
        lines.append((int(m.group(1), 16), cur))
src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
hd = src[1]
si, ii = hd.index("Warp Stall Sampling (All Samples)"), hd.index("Instructions Executed")
rows = [r for r in src[2:] if len(r) > ii]
print(f"# sass rows in report {len(rows)}, instructions in disassembly {len(lines)}")
agg = collections.defaultdict(lambda: [0, 0])
off2line = dict(lines)
a0 = int(rows[0][0], 16) if rows and rows[0][0].startswith("0x") else int(rows[0][0])
for r in rows:
    a = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
    l = off2line.get(a - a0)
    agg[l][0] += int(r[si] or 0); agg[l][1] += int(r[ii] or 0)
ts = sum(v[0] for v in agg.values()) or 1; ti = sum(v[1] for v in agg.values()) or 1
text = {}
for key_, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, n = key_ if key_ else ("?", 0)
    if f not in text:
        try:
            text[f] = open(f"morbit.jl_b200/csrc/{f}").read().splitlines()
        except OSError:
            text[f] = []
    code = text[f][n - 1].strip()[:110] if 0 < n <= len(text[f]) else ""
    print(f"{v[0] / ts * 100:5.1f}% samples {v[1] / ti * 100:5.1f}% instr  {f}:{n}  {code}")
