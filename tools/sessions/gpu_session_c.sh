#!/bin/bash
# One gpurun call: full GPU test suite, tall (C4) bench, tall reference arm.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_c.log
timeout 900 python bench.py --workload c4 --steps 2 --warmup 1 > gpurun_out/bench_c4_c.json 2> gpurun_out/bench_c4_c.err; echo "c4 rc=$?"
timeout 600 python bench.py --workload c4 --impl reference --steps 1 --warmup 1 > gpurun_out/bench_c4_ref_c.json 2> gpurun_out/bench_c4_ref_c.err; echo "c4ref rc=$?"
tail -5 gpurun_out/pytest_gpu_c.log; cat gpurun_out/bench_c4_c.json; tail -5 gpurun_out/bench_c4_c.err; cat gpurun_out/bench_c4_ref_c.json; tail -3 gpurun_out/bench_c4_ref_c.err
