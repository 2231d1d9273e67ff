#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for h in 0 1 2; do
  TCHEM_JIT_CACHE=0 TCHEM_JIT_L2HINT=$h timeout 300 python bench.py --workload C2 --extras none --no-e2e --no-cpu-baseline --steps 20 > gpurun_out/g_bench_h$h.json 2> gpurun_out/g_bench_h$h.err
  python - <<PY
import json
try:
    j=json.load(open('gpurun_out/g_bench_h$h.json'))
    print('L2 hint $h', '%.4g geom/s'%j['value'], 'ms %.4f'%j['ms_per_step'], 'frac %.3f'%j['roofline']['frac'], j['roofline']['kernels_launched'], j['jit'])
except Exception as e: print('hint $h bench failed', e); print(open('gpurun_out/g_bench_h$h.err').read()[-800:])
PY
done
