#!/bin/bash
# Builds tuning variants of libnsx.so that differ only in the TMA-fed SpMV's ring parameters
# (experiment helper; the chosen values become the defaults in kernels_spmv.cu).
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fopenmp,-O3 --expt-relaxed-constexpr"
OBJS=$(ls build/*.o | grep -v kernels_spmv)
for v in "$@"; do
  IFS=, read ts tnnz dl minb trows <<< "$v"
  tag="ts${ts}_n${tnnz}_dl${dl}_b${minb}_r${trows}"
  ( nvcc $FLAGS -DNSX_TS=$ts -DNSX_TNNZ=$tnnz -DNSX_DL=$dl -DNSX_TMINB=$minb -DNSX_TROWS=$trows -c navier_stokes_solver_b200/csrc/kernels_spmv.cu -o build/variants/spmv_$tag.o &&
    nvcc -shared -o build/variants/libnsx_$tag.so $OBJS build/variants/spmv_$tag.o -Xcompiler -fopenmp -lgomp -cudart shared && echo built $tag ) &
done
wait
