#!/bin/bash
# chained Cartesian kernel variants: parity tests, C5r bench line for bulk / copy-out / launch shapes, ncu summary of the default
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cartesian or chained" > gpurun_out/l_tests.log 2>&1
echo "rc $?" >> gpurun_out/l_tests.log
tail -n 4 gpurun_out/l_tests.log
for cfg in "1 1 8" "1 1 3" "1 2 1" "0 4 2"; do
  set -- $cfg
  echo "TCHEM_JIT_BULK=$1 TCHEM_JIT_CART_WARPS=$2 TCHEM_JIT_CART_BLOCKS=$3"
  TCHEM_JIT_BULK=$1 TCHEM_JIT_CART_WARPS=$2 TCHEM_JIT_CART_BLOCKS=$3 python bench.py --steps 10 --warmup 3 --workload C5r --extras none --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline'].get('frac'), d['roofline'].get('kernels_launched'))
"
done
