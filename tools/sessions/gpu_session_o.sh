#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/kbench.py --workload c2 > gpurun_out/kbench_o.log 2>&1; timeout 300 python tools/kbench.py --workload c3 --B 500000 >> gpurun_out/kbench_o.log 2>&1; cat gpurun_out/kbench_o.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_o.log; tail -4 gpurun_out/pytest_gpu_o.log
timeout 600 python bench.py --no-tall --no-cpu-baseline > gpurun_out/bench_c2_o.json 2> gpurun_out/bench_c2_o.err; echo "c2 rc=$?"; cut -c1-300 gpurun_out/bench_c2_o.json
