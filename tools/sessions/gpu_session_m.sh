#!/bin/bash
mkdir -p gpurun_out
BLSQ_GRAPH_DEBUG=1 timeout 600 python bench.py --no-tall --no-cpu-baseline --steps 2 --warmup 2 > gpurun_out/bench_c2_m.json 2> gpurun_out/bench_c2_m.err; echo "c2 rc=$?"; cut -c1-300 gpurun_out/bench_c2_m.json; grep "graph tail" gpurun_out/bench_c2_m.err | head -12
timeout 900 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/bench_c4_m.json 2> gpurun_out/bench_c4_m.err; echo "c4 rc=$?"; cut -c1-2500 gpurun_out/bench_c4_m.json
timeout 600 python -m pytest tests -m gpu -x -q -k "tall" > gpurun_out/pytest_gpu_m.log 2>&1; tail -3 gpurun_out/pytest_gpu_m.log
