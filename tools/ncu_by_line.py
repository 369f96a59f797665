"""Aggregate an .ncu-rep source page by CUDA source line: stall samples, executed instructions, shared wavefronts.
    python tools/ncu_by_line.py gpurun_out/prof_X.ncu-rep [file-substring] [top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; sub = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.OrderedDict(); cur = None; hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or cur is None or len(r) < len(hdr) or sub not in cur: continue
    try:
        ln = int(r[0])          # source-line rows carry the totals of their SASS rows (which have an empty line number)
    except ValueError:
        continue
    def g(name):
        try:
            return float(r[hdr.index(name)])
        except (ValueError, IndexError):
            return 0.0
    agg[(cur.split("/")[-1], ln)] = (g("Warp Stall Sampling (All Samples)"), g("Instructions Executed"), g("L1 Wavefronts Shared"), g("L1 Wavefronts Shared Excessive"), r[1].strip()[:110])
ts = sum(v[0] for v in agg.values()) or 1; ti = sum(v[1] for v in agg.values()) or 1; tw = sum(v[2] for v in agg.values()) or 1
print(f"total samples {ts:.0f} instr {ti:.0f} shared wavefronts {tw:.0f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:4d} smp {v[0]/ts*100:5.1f}% ins {v[1]/ti*100:5.1f}% wf {v[2]/tw*100:5.1f}% (exc {v[3]/tw*100:4.1f}%)  {v[4]}")
