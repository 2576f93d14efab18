#!/bin/bash
# n = 256 Gram kernels split over two CTAs: correctness (tall tests incl. n = 130) + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "tall" > gpurun_out/pytest_gpu_h1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_h1.log; tail -3 gpurun_out/pytest_gpu_h1.log
timeout 600 python tools/tall_check.py --time > gpurun_out/tall_check_h1.log 2>&1; grep -E '"n": (256|130|128)' gpurun_out/tall_check_h1.log | cut -c1-400
timeout 600 python bench.py --workload c5 --rows 2000000 --method trf --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c5small_trf_h1.json 2> gpurun_out/bench_c5small_trf_h1.err; echo "c5 trf rc=$?"; tail -2 gpurun_out/bench_c5small_trf_h1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c5small_trf_h1.json'))
print('c5small trf', d['value'], d['ms_per_step'], 'frac', d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['config']['iterations_per_step'], d['config']['status'])
PY
