#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_g2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_g2.log; grep -E "total|passed|failed|\(9, 'trf'\)" gpurun_out/pytest_gpu_g2.log | cut -c1-400 | tail -5
