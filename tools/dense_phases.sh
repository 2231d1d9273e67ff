#!/bin/bash
# Per-phase cycle counts of the dense per-geometry kernels (C4): builds an instrumented copy of the library
# (-DTCHEM_DENSE_PHASES, tools/scratch/, never the product), swaps it in for one run of tools/dense_phases.py
# and restores the product library. Step 1 (build) runs anywhere; step 2 needs the GPU:
#   bash tools/dense_phases.sh build        # here
#   bash tools/dense_phases.sh run          # on the GPU box (gpurun)
set -e
cd "$(dirname "$0")/.."
P=torch_chemistry_b200
case "$1" in
build)
  mkdir -p tools/scratch
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
       --expt-relaxed-constexpr -DTCHEM_DENSE_PHASES -x cu -c $P/csrc/ic_dense.cu -o tools/scratch/ic_dense_phases.o
  nvcc -shared -o tools/scratch/libtchem_b200_phases.so $(ls $P/build/*.o | grep -v ic_dense) tools/scratch/ic_dense_phases.o \
       -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -lcudart_static -lpthread -ldl -lrt
  ;;
run)
  cp $P/libtchem_b200.so /tmp/tchem_prod.so
  cp tools/scratch/libtchem_b200_phases.so $P/libtchem_b200.so
  python tools/dense_phases.py ${2:-4096} || true
  cp /tmp/tchem_prod.so $P/libtchem_b200.so
  ;;
*) echo "usage: $0 build|run [batch]"; exit 2;;
esac
