#!/bin/bash
# refresh of the transform-kernel evidence after a kernel change: phase cycle counts, ncu --set full of C4H and C4Hc (summaries made here)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
bash tools/sparse_phases.sh run > gpurun_out/i_phases.txt 2>&1
for spec in "C4H:hess_kernel" "C4Hc:factor_kernel|hess_kernel"; do
  w=${spec%%:*}; k=${spec#*:}
  n=1; skip=1; if [ "$w" = "C4Hc" ]; then n=2; skip=2; fi
  python tools/profile_all.py $w > /dev/null 2>&1 || exit 1
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:"$k" --launch-skip $skip -c $n -f -o gpurun_out/i_$w python tools/profile_all.py $w > gpurun_out/i_ncu_$w.log 2>&1
  echo "$w rc $?"
  python tools/ncu_summary.py gpurun_out/i_$w.ncu-rep 30 > gpurun_out/i_summary_$w.txt 2>&1
  rm -f gpurun_out/i_$w.ncu-rep
done
tail -n 32 gpurun_out/i_phases.txt
