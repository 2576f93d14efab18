#!/bin/bash
# final: full GPU suite + the C5 per-GPU shard (12.5M rows x 256) on one GPU with the split pass-1 kernel
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_h3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_h3.log; tail -3 gpurun_out/pytest_gpu_h3.log
timeout 600 python bench.py --workload c5 --rows 12500000 --method trf --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c5shard_trf_h3.json 2> gpurun_out/bench_c5shard_trf_h3.err; echo "c5 shard rc=$?"; tail -2 gpurun_out/bench_c5shard_trf_h3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c5shard_trf_h3.json'))
print('c5 shard trf', d['value'], d['ms_per_step'], 'frac', d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['config']['iterations_per_step'], d['config']['status'])
PY
