#!/bin/bash
# 3-point validation + per-round profile of C2/C3 + full ncu of the C3 (2-point) lin kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_i.log; tail -4 gpurun_out/pytest_gpu_i.log
timeout 600 python bench.py --no-tall --no-cpu-baseline --rounds-log gpurun_out/rounds_c2_i.csv > gpurun_out/bench_c2_i.json 2> gpurun_out/bench_c2_i.err; echo "c2 rc=$?"; cut -c1-600 gpurun_out/bench_c2_i.json
timeout 600 python bench.py --workload c3 --batch 2000000 --steps 2 --warmup 1 --no-cpu-baseline --rounds-log gpurun_out/rounds_c3_i.csv > gpurun_out/bench_c3_i.json 2> gpurun_out/bench_c3_i.err; echo "c3 rc=$?"; cut -c1-600 gpurun_out/bench_c3_i.json
timeout 300 python tools/kbench.py --workload c3 --B 500000 > gpurun_out/kbench_c3_i.log 2>&1; cat gpurun_out/kbench_c3_i.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lin_kernel -s 2 -c 1 -f -o gpurun_out/prof_lin_c3_r1 python tools/kbench.py --workload c3 --B 500000 --reps 2 > gpurun_out/ncu_lin_c3_i.log 2>&1; echo "ncu rc=$?"
