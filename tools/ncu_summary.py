"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + opcode mix + hottest SASS regions.
    python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep > profiles/ncu_X.txt"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit", "launch__grid_size", "launch__block_size",
        "smsp__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__pipe_tensor", "smsp__average_warp", "launch__shared_mem_per_block"]


def run(args):
    return subprocess.run(["ncu", "-i", sys.argv[1]] + args, capture_output=True, text=True).stdout


def main():
    raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
    h, units, vals = raw[0], raw[1], raw[2]
    print("== kernel:", vals[h.index("Kernel Name")] if "Kernel Name" in h else "?")
    for i, k in enumerate(h):
        if any(k.startswith(x) for x in KEYS):
            print(f"{k:85s} {vals[i]:>18s} {units[i]}")
    src = list(csv.reader(io.StringIO(run(["--page", "source", "--csv"]))))
    hd = src[1]
    si, ii = hd.index("Warp Stall Sampling (All Samples)"), hd.index("Instructions Executed")
    data = []
    for r in src[2:]:
        try:
            data.append((int(r[si] or 0), int(r[ii] or 0), r[1]))
        except (ValueError, IndexError):
            pass
    tot, toti = sum(d[0] for d in data) or 1, sum(d[1] for d in data) or 1
    op, ops = collections.Counter(), collections.Counter()
    for smp, ins, sass in data:
        parts = sass.split()
        o = (parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")).split(".")[0]
        op[o] += ins; ops[o] += smp
    print("\n== opcode mix (executed warp instructions / stall samples)")
    for o, c in op.most_common(16):
        print(f"{o:12s} instr {c / toti * 100:5.1f}%   samples {ops[o] / tot * 100:5.1f}%")
    print("\n== hottest SASS instructions by stall samples")
    for smp, ins, sass in sorted(data, reverse=True)[:14]:
        print(f"{smp / tot * 100:5.1f}% samples {ins / toti * 100:5.2f}% instr   {sass[:100]}")


if __name__ == "__main__":
    main()
