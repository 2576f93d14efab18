#!/bin/bash
# tall: tests, round latency, C4 bench, ncu launch list + full capture of the Gram kernels
mkdir -p gpurun_out
timeout 200 python tools/tall_check.py --round 2>&1 | tail -4
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_f.log 2>&1; tail -3 gpurun_out/pytest_gpu_f.log
timeout 600 python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/bench_c4_f.json 2> gpurun_out/bench_c4_f.err; cat gpurun_out/bench_c4_f.json; tail -3 gpurun_out/bench_c4_f.err
BENCH="python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 $BENCH > gpurun_out/plain_f.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_c4_r1.csv $BENCH > gpurun_out/ncu_launch_f.log 2>&1
timeout 300 $BENCH > gpurun_out/plain_f2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 4 -c 4 -f -o gpurun_out/prof_gram_r1 $BENCH > gpurun_out/ncu_gram_f.log 2>&1
tail -2 gpurun_out/ncu_launch_f.log gpurun_out/ncu_gram_f.log
