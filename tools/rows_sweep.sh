#!/bin/bash
for rows in 2 3 4; do
  for w in C3 C4J; do
    TCHEM_GROUP_ROWS=$rows TCHEM_DEBUG=1 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/e.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rows=$rows $w', round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"
    grep "group kernel" gpurun_out/e.err | tail -1
  done
done
