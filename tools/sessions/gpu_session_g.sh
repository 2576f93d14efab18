#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/tall_check.py --quick 2>&1 | tail -1 | cut -c1-420
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_g.log 2>&1; tail -3 gpurun_out/pytest_gpu_g.log
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4_g.json 2> gpurun_out/bench_c4_g.err; cat gpurun_out/bench_c4_g.json; tail -3 gpurun_out/bench_c4_g.err
timeout 200 python tools/tall_check.py --round > gpurun_out/round_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tall_round_kernel -s 3 -c 2 -f -o gpurun_out/prof_round_tall_r1 python tools/tall_check.py --round > gpurun_out/ncu_round_g.log 2>&1
tail -n 2 gpurun_out/ncu_round_g.log
