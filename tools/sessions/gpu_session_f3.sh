#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_f3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_f3.log; tail -3 gpurun_out/pytest_gpu_f3.log
timeout 600 python bench.py --workload c3 --batch 2000000 --steps 3 --warmup 1 --no-cpu-baseline --rounds-log gpurun_out/rounds_c3_f3.csv > gpurun_out/bench_c3_f3.json 2> gpurun_out/bench_c3_f3.err; echo "c3 rc=$?"; tail -2 gpurun_out/bench_c3_f3.err
python - <<'PY'
import json
for f in ('bench_c3_f3',):
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, 'value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'cb ms', d['roofline']['callbacks_ms_per_step'], 'kern ms', d['roofline']['kernels_ms_per_step'])
PY
head -12 gpurun_out/rounds_c3_f3.csv
