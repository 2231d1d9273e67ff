#!/bin/bash
# C2 q+J through the bulk-store (TMA write-out) specialised kernel at several launch shapes / chunk sizes
for cfg in "2 4 76" "4 2 76" "1 8 76" "2 5 76" "2 6 76" "2 5 38" "2 6 38" "2 4 38" "2 3 114" "3 2 114" "2 4 57"; do
  set -- $cfg
  echo "bulk warps=$1 blocks=$2 dcap=$3: $(TCHEM_JIT_WARPS=$1 TCHEM_JIT_BLOCKS=$2 TCHEM_JIT_DCAP=$3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.4f ms  %.3f" % (d["ms_per_step"], d["roofline"]["frac"]))')"
done
