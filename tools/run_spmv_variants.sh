#!/bin/bash
# GPU side of the SpMV tuning experiment: correctness of each variant, then CUDA-event timings.
for so in build/variants/libnsx_*.so; do
  echo "== $so"
  NSX_LIB=$PWD/$so python -m pytest tests/test_gpu_linalg.py -q -m gpu -k "spmv" 2>&1 | tail -1
  NSX_LIB=$PWD/$so python tools/profile_step.py --kernels-only --kernels 0,1 --reps 50 --spmv 3 --mode 1 2>&1 | grep "kernel [0-9]"
done
