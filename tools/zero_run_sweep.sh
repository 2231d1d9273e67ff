#!/bin/bash
for z in 64 32 16 8; do
  for w in C3 C3K C2; do
    extra=""; [ $w = C3K ] && extra="--batch 8192"
    [ $w = C2 ] && export TCHEM_JIT=0 || unset TCHEM_JIT
    TCHEM_MIN_ZERO_RUN=$z python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e $extra 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('zrun=$z $w', round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"
  done
done
