#!/usr/bin/env python
"""Stall samples per CUDA source line of ONE launch of an .ncu-rep, grouped into phases.
usage: python tools/ncu_lines.py rep launch_index [top_n]"""
import collections, csv, subprocess, sys
rep, launch = sys.argv[1], int(sys.argv[2])
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--launch-skip", str(launch), "--launch-count", "1"], capture_output=True, text=True).stdout
cur, h, agg, tot = None, None, collections.OrderedDict(), 0
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] in ("File Name", "File Path"): cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": h = r; iS = h.index("# Samples"); iI = h.index("Instructions Executed"); continue
    if h is None or not r[0].isdigit() or len(r) <= iI: continue
    try: s, n = int(float(r[iS] or 0)), int(float(r[iI] or 0))
    except ValueError: continue
    key = (cur, int(r[0]))
    if key in agg: continue           # the cuda view comes first; sass rows repeat line numbers as addresses
    agg[key] = (s, n, r[1][:90]); tot += s
print("total samples", tot)
for (f, l), (s, n, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print("%6d %5.1f%%  %-14s L%-4d inst %10d  %s" % (s, 100.0 * s / max(tot, 1), f, l, n, src))
