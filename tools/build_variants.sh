#!/bin/bash
# Builds tuning variants of the batched kernels into tools/variants/ (N=4,6 only).
# usage: tools/build_variants.sh name "-DFLAG=.. -DFLAG=.." [name flags]...
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/variants
C=bounded_lsq_b200/csrc
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
     -Xcompiler -fPIC -Xptxas -v -DBLSQ_ONLY_N46 $flags -shared \
     $C/blsq_batched.cu $C/blsq_elementwise.cu -o tools/variants/lib_$name.so -lcudart \
     2> tools/variants/$name.ptxas.log && echo built $name) &
done
wait
