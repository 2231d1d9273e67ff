#!/bin/bash
# validation session (one GPU): whole suite, smoke, default bench (all extras), reference arm, launch lists
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/f_full.log 2>&1
echo "rc $?" >> gpurun_out/f_full.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1
echo "rc $?" >> gpurun_out/f_smoke.log
( time timeout 1500 python bench.py > gpurun_out/f_bench_default.json 2> gpurun_out/f_bench_default.err ) 2> gpurun_out/f_bench_time.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_reference.json 2> gpurun_out/f_bench_reference.err
# launch lists: the headline alone (what the timed region of `value` launches), then the extras without the host-buffer legs
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches_headline.csv \
    python bench.py --steps 20 --warmup 5 --extras none --no-e2e --no-cpu-baseline > gpurun_out/f_ncu_headline.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/f_launches_extras.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/f_ncu_extras.log 2>&1
for f in f_full f_smoke; do tail -n 4 gpurun_out/$f.log; done; cat gpurun_out/f_bench_time.txt
