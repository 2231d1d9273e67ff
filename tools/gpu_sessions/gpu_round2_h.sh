#!/bin/bash
# quick check of a transform-kernel change: transform parity tests, C4 extras, phase probe
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "transform or hessian or gradient or cart2int or dense or alternating or smoke" > gpurun_out/h_tests.log 2>&1
echo "rc $?" >> gpurun_out/h_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --extras C4H,C4Hc,C4gc,C4g --no-e2e --no-cpu-baseline > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
bash tools/sparse_phases.sh run > gpurun_out/h_phases.txt 2>&1
tail -n 5 gpurun_out/h_tests.log; cat gpurun_out/h_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d.get('extra',{}).items(): print(k, v.get('value'), v.get('ms_per_step'))
"; tail -n 30 gpurun_out/h_phases.txt
