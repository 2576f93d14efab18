#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_q.log; tail -4 gpurun_out/pytest_gpu_q.log
timeout 600 python bench.py --no-tall --no-cpu-baseline --rounds-log gpurun_out/rounds_c2_q.csv > gpurun_out/bench_c2_q.json 2> gpurun_out/bench_c2_q.err; echo "c2 rc=$?"; cut -c1-300 gpurun_out/bench_c2_q.json; tail -3 gpurun_out/bench_c2_q.err
timeout 600 python bench.py --workload c3 --batch 2000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c3_q.json 2> gpurun_out/bench_c3_q.err; echo "c3 rc=$?"; cut -c1-300 gpurun_out/bench_c3_q.json; tail -3 gpurun_out/bench_c3_q.err
python - <<'PY'
import json
for f in ('bench_c2_q','bench_c3_q'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, 'value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'])
    except Exception as e: print(f, e)
PY
