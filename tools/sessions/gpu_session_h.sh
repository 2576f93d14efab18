#!/bin/bash
# round-end style validation: GPU tests, smoke, default bench (c2 + tall), reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_h.log; tail -4 gpurun_out/pytest_gpu_h.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_h.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke_h.log
timeout 900 python bench.py > gpurun_out/bench_default_h.json 2> gpurun_out/bench_default_h.err; echo "bench rc=$?"; cut -c1-1200 gpurun_out/bench_default_h.json; tail -3 gpurun_out/bench_default_h.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_h.json 2> gpurun_out/bench_ref_h.err; echo "ref rc=$?"; cut -c1-1500 gpurun_out/bench_ref_h.json; tail -3 gpurun_out/bench_ref_h.err
