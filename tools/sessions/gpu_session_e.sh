#!/bin/bash
mkdir -p gpurun_out
for v in g h i; do
  BLSQ_B200_LIB=$PWD/tools/variants/libg_$v.so timeout 120 python tools/tall_check.py --quick 2>&1 | tail -1
done > gpurun_out/gram_variants2.log 2>&1
cut -c1-420 gpurun_out/gram_variants2.log
timeout 120 python tools/tall_check.py --quick 2>&1 | tail -1 | cut -c1-420
timeout 200 python tools/tall_check.py --round 2>&1 | tail -5
timeout 600 python -m pytest tests -m gpu -x -q -k tall > gpurun_out/pytest_tall3.log 2>&1; tail -3 gpurun_out/pytest_tall3.log
timeout 600 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c4_e.json 2> gpurun_out/bench_c4_e.err; cat gpurun_out/bench_c4_e.json; tail -3 gpurun_out/bench_c4_e.err
