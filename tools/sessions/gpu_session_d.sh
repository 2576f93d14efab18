#!/bin/bash
# gram kernel variant sweep (n = 64) + ncu of the default build's gram kernels
mkdir -p gpurun_out
for v in a b c d e f; do
  BLSQ_B200_LIB=$PWD/tools/variants/libg_$v.so timeout 120 python tools/tall_check.py --quick 2>&1 | tail -3
done > gpurun_out/gram_variants.log 2>&1
cat gpurun_out/gram_variants.log | cut -c1-600
