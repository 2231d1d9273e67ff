#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
python tools/dense_once.py 2048 > gpurun_out/c_once.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"hess_kernel|factor_kernel" --launch-skip 6 -c 6 -f -o gpurun_out/c_sparse python tools/dense_once.py 2048 > gpurun_out/c_ncu.log 2>&1
tail -3 gpurun_out/c_once.log; tail -3 gpurun_out/c_ncu.log; ls -la gpurun_out/c_sparse.ncu-rep
