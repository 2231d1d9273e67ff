#!/bin/bash
# every workload once, kernel numbers only (no CPU baseline / e2e): one JSON line per workload
mkdir -p gpurun_out
: > gpurun_out/bench_all.jsonl
for w in C1 C2 C3 C3K C4J C4g C4gc C4H C4Hc C5 C5x C5K; do
  extra=""
  case $w in C3K) extra="--batch 8192";; C4H|C4Hc|C4gc) extra="--batch 20000";; esac
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e $extra 2>gpurun_out/bench_$w.err | tail -1 >> gpurun_out/bench_all.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/bench_all.jsonl'):
    try: d = json.loads(l)
    except Exception: print('BAD', l[:200]); continue
    r = d['roofline']
    print("%-60s %12.4g geom/s  %8.3f ms  %s %.4g %s  frac %.3f" % (d['config']['workload'][:60], d['value'], d['ms_per_step'], r['bound'], r['achieved'], r['unit'], r['frac']))
PY
