#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_g5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_g5.log; tail -4 gpurun_out/pytest_gpu_g5.log
