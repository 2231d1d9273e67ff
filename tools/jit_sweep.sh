#!/bin/bash
# C2 q+J through the specialised kernel at several launch shapes / chunk sizes
for cfg in "4 2 76" "3 3 76" "2 4 76" "2 5 57" "3 3 57" "5 2 57" "1 9 76"; do
  set -- $cfg
  echo "warps=$1 blocks=$2 dcap=$3: $(TCHEM_JIT_WARPS=$1 TCHEM_JIT_BLOCKS=$2 TCHEM_JIT_DCAP=$3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.4f ms  %.3f" % (d["ms_per_step"], d["roofline"]["frac"]))')"
done
echo "interpreter: $(TCHEM_JIT=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.4f ms  %.3f" % (d["ms_per_step"], d["roofline"]["frac"]))')"
