#!/bin/bash
# One gpurun call: FP64 peaks, GPU tests, bench, launch list, full ncu of the two batched kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_a.csv
./tools/fp64_peak > gpurun_out/fp64_peak.jsonl 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_a.log
python bench.py > gpurun_out/bench_c2_a.json 2> gpurun_out/bench_c2_a.err
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$BENCH > gpurun_out/plain_a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1b.csv $BENCH > gpurun_out/ncu_launch_a.log 2>&1
$BENCH > gpurun_out/plain_a2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lin_kernel -s 2 -c 2 -f -o gpurun_out/prof_lin_r1b $BENCH > gpurun_out/ncu_lin_a.log 2>&1
$BENCH > gpurun_out/plain_a3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:round_kernel -s 2 -c 2 -f -o gpurun_out/prof_round_r1b $BENCH > gpurun_out/ncu_round_a.log 2>&1
tail -3 gpurun_out/pytest_gpu_a.log; cat gpurun_out/fp64_peak.jsonl; cat gpurun_out/bench_c2_a.json
