#!/bin/bash
# graph-tail validation + C2 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "graph_tail or compaction or golden_exact" > gpurun_out/pytest_gpu_j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_j.log; tail -12 gpurun_out/pytest_gpu_j.log
timeout 600 python bench.py --no-tall --no-cpu-baseline > gpurun_out/bench_c2_j.json 2> gpurun_out/bench_c2_j.err; echo "c2 rc=$?"; cut -c1-900 gpurun_out/bench_c2_j.json; tail -5 gpurun_out/bench_c2_j.err
timeout 600 python bench.py --no-tall --no-cpu-baseline --callbacks torch > gpurun_out/bench_c2_torch_j.json 2> gpurun_out/bench_c2_torch_j.err; echo "c2 torch rc=$?"; cut -c1-400 gpurun_out/bench_c2_torch_j.json; tail -5 gpurun_out/bench_c2_torch_j.err
