#!/usr/bin/env python
"""SASS evidence for profiles/rNN_sass_summary.txt: instruction counts per kernel of the built library (and of NVRTC
cubins from the on-disk JIT cache of a GPU session, if a directory of them is given).

    python tools/sass_summary.py [cubin_dir] > profiles/r02_sass_summary.txt

Runs without a GPU (cuobjdump only)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MNEMONICS = ["DMMA", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "ATOMS", "ATOMG", "DFMA", "DMUL", "DADD", "LDS", "STS", "SHFL", "MUFU",
             "STG", "LDG", "STL", "LDL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"tchem_b200::\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\((tchem_b200::)?SparseProgram.*$", "", name)
    name = re.sub(r"\(tchem_.*$", "", name)
    return name[:60]


def summarise(path):
    res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+)", line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)))
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,8}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1).split(".")[0]
            counts[cur]["instr"] += 1
            if op in MNEMONICS:
                counts[cur][op] += 1
    names = demangle(list(counts))
    rows = []
    for fn, c in counts.items():
        regs, stack = usage.get(fn, (0, 0))
        text = "%-60s regs %3d (stack %d B)%s instr %6d  " % (short(names.get(fn, fn)), regs, stack, " " * max(1, 8 - len(str(stack))), c["instr"])
        text += "  ".join("%s %d" % (k, c[k]) for k in MNEMONICS if c[k])
        rows.append((c["instr"], text))
    return [t for _, t in sorted(rows, reverse=True)]


def main():
    print("DMMA = mma.sync.m8n8k4.f64 (FP64 tensor core), UTMASTG = cp.async.bulk.tensor store (TMA), UBLKCP = cp.async.bulk 1-D copy (TMA,")
    print("loads and stores), SYNCS = mbarrier operations, LDGSTS = cp.async (global -> shared without registers), ATOMS / ATOMG = shared /")
    print("global atomics, STL / LDL = local-memory spills\n")
    lib = os.path.join(ROOT, "torch_chemistry_b200", "libtchem_b200.so")
    print("== torch_chemistry_b200/libtchem_b200.so (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo): SASS instruction counts per "
          "kernel (cuobjdump -sass / -res-usage; tools/sass_summary.py)")
    for t in summarise(lib):
        print(t)
    if len(sys.argv) > 1 and os.path.isdir(sys.argv[1]):
        for f in sorted(os.listdir(sys.argv[1])):
            if f.endswith(".cubin"):
                print("== NVRTC cubin (sm_100a) from the on-disk JIT cache of a GPU session: " + f)
                for t in summarise(os.path.join(sys.argv[1], f)):
                    print(t)


if __name__ == "__main__":
    main()
