#!/usr/bin/env python
"""Summarise an .ncu-rep (read here on the CPU box): key metrics + instruction share per source line.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__cycles_elapsed.max",
        "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_lsu.sum", "smsp__inst_executed_pipe_alu.sum",
        "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_cbu.sum",
        "smsp__inst_executed_pipe_uniform.sum", "smsp__inst_executed_pipe_adu.sum"]
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print("%-75s %-14s %s" % (w, units[i], [r[i] for r in data]))
print("-- stall samples (launch 0)")
for i, h in enumerate(hdr):
    if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h:
        print("   %-30s %s" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), data[0][i]))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = collections.OrderedDict()
tot = 0
cur = None
h = None
launch = -1
for r in csv.reader(src.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        h = r
        iI, iS = h.index("Instructions Executed"), h.index("# Samples")
        continue
    if h is None or len(r) <= iI:
        continue
    try:
        inst, samp = int(float(r[iI] or 0)), int(float(r[iS] or 0))
    except ValueError:
        continue
    if not r[0] or not r[0].isdigit():
        continue                      # SASS-view rows repeat the same counts
    key = (cur, r[0], r[1].strip()[:100])
    a = agg.setdefault(key, [0, 0])
    a[0] += inst
    a[1] += samp
    tot += inst
print("-- instructions per source line (all profiled launches summed): total", tot)
for (f, ln, s), (inst, samp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print("%-14s%5s %12d %5.1f%% samples %6d  %s" % (f, ln, inst, 100.0 * inst / max(tot, 1), samp, s))

# ---- top lines by stall samples with the reasons
import re
agg2 = collections.OrderedDict()
cur = None
h = None
for r in csv.reader(src.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        h = r
        stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
        iS = h.index("# Samples")
        continue
    if h is None or not r[0].isdigit() or len(r) <= iS:
        continue
    key = (cur, r[0], r[1].strip()[:70])
    d = agg2.setdefault(key, collections.Counter())
    try:
        d["#"] += int(float(r[iS] or 0))
        for i, n in stall_cols:
            if i < len(r) and r[i]:
                d[n] += int(float(r[i]))
    except ValueError:
        pass
print("-- top lines by stall samples")
for (f, ln, s_), d in sorted(agg2.items(), key=lambda kv: -kv[1]["#"])[:25]:
    reasons = ", ".join("%s %d" % (k.replace("stall_", ""), v) for k, v in d.most_common(4) if k != "#")
    print("%-14s%5s %6d  %-70s | %s" % (f, ln, d["#"], s_, reasons))
