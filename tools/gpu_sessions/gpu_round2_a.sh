#!/bin/bash
# GPU session A of round 2: new sparse-J kernels vs the oracle, then the whole suite, smoke and the bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "transform or hessian or cart2int or dense or alternating or round_trip" > gpurun_out/a_t1.log 2>&1
echo "rc $?" >> gpurun_out/a_t1.log
for mode in 0 1; do
  TCHEM_DENSE_DMMA=$mode timeout 600 python bench.py --workload C4gc --extras C4H,C4Hc,C4g --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/a_bench_dense_mode$mode.json 2> gpurun_out/a_bench_dense_mode$mode.err
done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/a_full.log 2>&1
echo "rc $?" >> gpurun_out/a_full.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/a_smoke.log 2>&1
echo "rc $?" >> gpurun_out/a_smoke.log
timeout 1200 python bench.py > gpurun_out/a_bench_default.json 2> gpurun_out/a_bench_default.err
echo "rc $?" >> gpurun_out/a_bench_default.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/a_bench_reference.json 2> gpurun_out/a_bench_reference.err
for f in a_t1 a_full a_smoke; do tail -n 3 gpurun_out/$f.log; done
