#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/funbench_v2.log
for f in tools/variants/libmodels_u16.so tools/variants/libmodels_v1rb4.so tools/variants/libmodels_v1rb8.so tools/variants/libmodels_v1rb16.so; do
  timeout 120 python tools/funbench.py $f >> gpurun_out/funbench_v2.log 2>&1; echo "$f rc=$?" >> gpurun_out/funbench_v2.log
done
cat gpurun_out/funbench_v2.log | cut -c1-300
