#!/bin/bash
# Per-phase cycle counts of the sparse-J transform kernels (C4): builds an instrumented copy of the library
# (-DTCHEM_SPARSE_PHASES, tools/scratch/, never the product), swaps it in for one run of tools/sparse_phases.py and
# restores the product library.   bash tools/sparse_phases.sh build   (here);   bash tools/sparse_phases.sh run   (GPU box)
set -e
cd "$(dirname "$0")/.."
P=torch_chemistry_b200
case "$1" in
build)
  mkdir -p tools/scratch
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
       --expt-relaxed-constexpr -DTCHEM_SPARSE_PHASES -x cu -c $P/csrc/ic_sparse.cu -o tools/scratch/ic_sparse_phases.o
  nvcc -shared -o tools/scratch/libtchem_b200_sphases.so $(ls $P/build/*.o | grep -v ic_sparse) tools/scratch/ic_sparse_phases.o \
       -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -lcudart_static -lpthread -ldl -lrt
  ;;
run)
  cp $P/libtchem_b200.so /tmp/tchem_prod.so
  cp tools/scratch/libtchem_b200_sphases.so $P/libtchem_b200.so
  python tools/sparse_phases.py ${2:-8192} || true
  cp /tmp/tchem_prod.so $P/libtchem_b200.so
  ;;
*) echo "usage: $0 build|run [batch]"; exit 2;;
esac
