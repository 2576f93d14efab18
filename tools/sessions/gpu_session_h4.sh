#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_h4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_h4.log; tail -3 gpurun_out/pytest_gpu_h4.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_h4.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_h4.log
