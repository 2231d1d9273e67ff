#!/bin/bash
for w in C3 C3K C2i C1 C5i; do
  extra=""; wl=$w
  [ $w = C3K ] && extra="--batch 8192"
  unset TCHEM_JIT
  [ $w = C2i ] && { export TCHEM_JIT=0; wl=C2; }
  [ $w = C5i ] && { export TCHEM_JIT=0; wl=C5; }
  python bench.py --workload $wl --steps 10 --warmup 5 --no-cpu-baseline --no-e2e $extra 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"
done
