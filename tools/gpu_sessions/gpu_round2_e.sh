#!/bin/bash
# profile session of round 2 (one GPU): launch list of the default bench, one ncu --set full capture per serving kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python tools/profile_all.py > gpurun_out/e_profile_all.log 2>&1 || exit 1
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/e_launch_list.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench_under_ncu.log 2>&1
for spec in "C2:tchem_jit_qJ_bulk" "C3:ic_eval_group_kernel" "C3K:ic_eval_kernel" "C4g:grad_int2cart" "C4gc:factor_kernel" "C4H:hess_kernel" \
            "C4Hc:factor_kernel|hess_kernel" "C5:tchem_jit_chain_bulk" "C5K:tchem_jit_sap_hess" "C5r:tchem_jit_chain_cart"; do
  w=${spec%%:*}; k=${spec#*:}
  n=1; skip=1; if [ "$w" = "C4Hc" ]; then n=2; skip=2; fi
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:"$k" --launch-skip $skip -c $n -f -o gpurun_out/e_$w python tools/profile_all.py $w > gpurun_out/e_ncu_$w.log 2>&1
  echo "$w rc $?"
  # summaries are made here (the reports together exceed what travels back); the large reports are dropped
  python tools/ncu_summary.py gpurun_out/e_$w.ncu-rep 30 > gpurun_out/e_summary_$w.txt 2>&1
  case $w in C3|C3K|C4gc|C4Hc) rm -f gpurun_out/e_$w.ncu-rep;; esac
done
mkdir -p gpurun_out/e_cubins; cp ~/.cache/tchem_b200/*.cubin gpurun_out/e_cubins/ 2>/dev/null; ls gpurun_out/e_cubins | wc -l
ls -la gpurun_out/e_*.ncu-rep
