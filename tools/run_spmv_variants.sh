#!/bin/bash
# GPU side of the SpMV tuning experiment: correctness of each variant, then CUDA-event timings (flushed + back to back).
for so in build/variants/libnsx_*.so; do
  echo "== $so"
  NSX_LIB=$PWD/$so timeout 120 python -m pytest tests/test_gpu_linalg.py -q -m gpu -k "spmv" 2>&1 | tail -1
  NSX_LIB=$PWD/$so timeout 120 python tools/spmv_timing.py 2>&1 | grep "^mode 3"
done
