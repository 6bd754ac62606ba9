#!/usr/bin/env python3
"""Summarise an Nsight Compute report of the tracker kernel as markdown (run here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" [stages_per_launch] >> profiles/ncu_r1.md
Key raw metrics, warp-stall breakdown per issued instruction, and executed instructions per HC stage by opcode."""
import collections
import csv
import subprocess
import sys

rep, title = sys.argv[1], sys.argv[2]
stages = float(sys.argv[3]) if len(sys.argv) > 3 else 8811960.0      # default round: 31 200 paths x 282.43 stages

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
d = dict(zip(rows[0], rows[2]))
u = dict(zip(rows[0], rows[1]))
f = lambda k: float(d[k].replace(",", ""))
print("### %s\n" % title)
print("| metric | value |\n|---|---|")
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__icc_request_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum"]
for k in keys:
    if k in d:
        print("| `%s` | %s %s |" % (k, d[k], u[k]))
print("| executed warp-instructions per HC stage | %.0f |" % (f("smsp__inst_executed.sum") / stages))
print("\nWarp stalls per issued instruction (`smsp__average_warps_issue_stalled_*_per_issue_active`, > 0.05):\n")
st = []
for h in rows[0]:
    if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
        v = f(h)
        if v > 0.05:
            st.append((v, h.split("stalled_")[1].split("_per_issue")[0]))
print(", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, data = rows[1], rows[2:]
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
h, hs = collections.Counter(), collections.Counter()
tot_s = 0
for r in data:
    t = r[iS].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    h[op] += int(r[iE]); hs[op] += int(r[iN]); tot_s += int(r[iN])
print("\nExecuted instructions per stage by opcode (and share of stall samples), %d SASS instructions in the kernel:\n" % len(data))
print(", ".join("%s %.0f (%.0f%%)" % (o, c / stages, 100.0 * hs[o] / max(tot_s, 1)) for o, c in h.most_common(18)))
print()
