#!/bin/bash
for line in 1 0; do
for g in 0 4 6; do
  for w in C3 C4J; do
    TCHEM_LINE=$line TCHEM_DEBUG=1 TCHEM_GROUP_G=$g python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/grp.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('line=$line G=$g $w', round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"
    grep "ic_eval" gpurun_out/grp.err | tail -1
  done
done
done
TCHEM_JIT=0 python bench.py --workload C2 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C2 interp', round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"
