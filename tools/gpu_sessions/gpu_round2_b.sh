#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "transform or hessian or cart2int or dense or alternating or round_trip or 32_reference" > gpurun_out/b_t1.log 2>&1
echo "rc $?" >> gpurun_out/b_t1.log
timeout 600 bash tools/sparse_phases.sh run 8192 > gpurun_out/b_phases.txt 2>&1
timeout 600 python bench.py --workload C4gc --extras C4H,C4Hc --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
tail -3 gpurun_out/b_t1.log; cat gpurun_out/b_phases.txt; tail -3 gpurun_out/b_bench.err; python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/b_bench.json'))
    recs={'C4gc':j}; recs.update(j.get('extra',{}))
    for k,r in recs.items(): print(k, '%.3g geom/s'%r['value'], 'ms %.3f'%r['ms_per_step'], r['roofline']['kernels_launched'])
except Exception as e: print('bench parse failed', e)
PY
