#!/bin/bash
for k in 2 3 4; do
  TCHEM_K_WARPS=$k TCHEM_DEBUG=1 python bench.py --workload C3K --batch 8192 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/e.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('kwarps=$k C3K', round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"
  grep "ic_eval mode 2" gpurun_out/e.err | tail -1
done
