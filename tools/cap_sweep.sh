#!/bin/bash
for z in 0 64 84 112 168 224; do
  for w in C3; do
    TCHEM_DEBUG=1 TCHEM_CAP_MIN=$z python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/cap.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('capmin=$z $w', round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"
    grep "ic_eval mode" gpurun_out/cap.err | tail -1
  done
done
