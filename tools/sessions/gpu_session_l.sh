#!/bin/bash
# manual graph capture (no allocator flush), odd n in tall mode
mkdir -p gpurun_out
timeout 1500 python -W always -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_l.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_l.log; grep -i "captured" gpurun_out/pytest_gpu_l.log | head -4; tail -4 gpurun_out/pytest_gpu_l.log
timeout 600 python -W always bench.py --no-tall --no-cpu-baseline --rounds-log gpurun_out/rounds_c2_l.csv > gpurun_out/bench_c2_l.json 2> gpurun_out/bench_c2_l.err; echo "c2 rc=$?"; cut -c1-300 gpurun_out/bench_c2_l.json; grep -i "captured" gpurun_out/bench_c2_l.err | head -3
BLSQ_GRAPH_TAIL=0 timeout 600 python bench.py --no-tall --no-cpu-baseline > gpurun_out/bench_c2_l_nograph.json 2> gpurun_out/bench_c2_l_nograph.err; echo "c2 nograph rc=$?"; cut -c1-300 gpurun_out/bench_c2_l_nograph.json
timeout 600 python bench.py --no-tall --no-cpu-baseline --callbacks torch > gpurun_out/bench_c2_torch_l.json 2> gpurun_out/bench_c2_torch_l.err; echo "c2 torch rc=$?"; cut -c1-300 gpurun_out/bench_c2_torch_l.json
timeout 600 python bench.py --workload c3 --batch 2000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c3_l.json 2> gpurun_out/bench_c3_l.err; echo "c3 rc=$?"; cut -c1-300 gpurun_out/bench_c3_l.json
