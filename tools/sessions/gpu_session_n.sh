#!/bin/bash
mkdir -p gpurun_out
BLSQ_GRAPH_DEBUG=1 timeout 600 python bench.py --no-tall --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/bench_c2_n.json 2> gpurun_out/bench_c2_n.err; echo "c2 rc=$?"; cut -c1-300 gpurun_out/bench_c2_n.json; grep "graph tail\|Warning" gpurun_out/bench_c2_n.err | head -12
BLSQ_GRAPH_DEBUG=1 timeout 600 python bench.py --workload c3 --batch 2000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c3_n.json 2> gpurun_out/bench_c3_n.err; echo "c3 rc=$?"; cut -c1-300 gpurun_out/bench_c3_n.json;  grep "graph tail\|Warning" gpurun_out/bench_c3_n.err | head -6
