#!/bin/bash
# ncu --set full of one workload's serving kernel (summary made here): bash tools/gpu_sessions/gpu_round2_m.sh C5r tchem_jit_chain_cart
cd "$(dirname "$0")/../.."
w=${1:-C5r}; k=${2:-tchem_jit_chain_cart}
mkdir -p gpurun_out
python tools/profile_all.py $w > /dev/null 2>&1 || exit 1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"$k" --launch-skip 1 -c 1 -f -o gpurun_out/m_$w python tools/profile_all.py $w > gpurun_out/m_ncu_$w.log 2>&1
echo "$w rc $?"
python tools/ncu_summary.py gpurun_out/m_$w.ncu-rep 30 > gpurun_out/m_summary_$w.txt 2>&1
rm -f gpurun_out/m_$w.ncu-rep
mkdir -p gpurun_out/m_cubins; cp ~/.cache/tchem_b200/*.cubin gpurun_out/m_cubins/ 2>/dev/null
head -30 gpurun_out/m_summary_$w.txt
