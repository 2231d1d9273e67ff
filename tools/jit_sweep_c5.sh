#!/bin/bash
for cfg in "4 2" "2 4" "2 5" "3 3" "4 3" "2 6"; do
  set -- $cfg
  for w in C5 C5x; do
  echo "$w warps=$1 blocks=$2: $(TCHEM_JIT_WARPS=$1 TCHEM_JIT_BLOCKS=$2 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.4f ms  %.3f" % (d["ms_per_step"], d["roofline"]["frac"]))')"
  done
done
