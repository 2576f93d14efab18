#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/funbench.py > gpurun_out/funbench_v.log 2>&1; cat gpurun_out/funbench_v.log
