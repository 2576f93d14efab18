#!/bin/bash
# new lin kernel (RPL 8 for n<=6, rsqrt norms) + graph-tail diagnosis
mkdir -p gpurun_out
timeout 300 python tools/kbench.py --workload c2 > gpurun_out/kbench_k.log 2>&1; timeout 300 python tools/kbench.py --workload c3 --B 500000 >> gpurun_out/kbench_k.log 2>&1; cat gpurun_out/kbench_k.log
timeout 1500 python -W always -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_k.log; grep -i "captured\|Warning" gpurun_out/pytest_gpu_k.log | head -8; tail -4 gpurun_out/pytest_gpu_k.log
timeout 600 python -W always bench.py --no-tall --no-cpu-baseline > gpurun_out/bench_c2_k.json 2> gpurun_out/bench_c2_k.err; echo "c2 rc=$?"; cut -c1-300 gpurun_out/bench_c2_k.json; grep -i "captured" gpurun_out/bench_c2_k.err | head -3
timeout 600 python bench.py --workload c3 --batch 2000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c3_k.json 2> gpurun_out/bench_c3_k.err; echo "c3 rc=$?"; cut -c1-300 gpurun_out/bench_c3_k.json
