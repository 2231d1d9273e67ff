#!/bin/bash
for g in 4 6; do
for rows in 2 3 4; do
  for w in C3 C4J; do
    TCHEM_GROUP_ROWS=$rows TCHEM_DEBUG=1 TCHEM_GROUP_G=$g python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/grp.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('G=$g rows=$rows $w', round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"
    grep "group kernel" gpurun_out/grp.err | tail -1
  done
done
done
