#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --no-tall > gpurun_out/bench_c2_f8.json 2> gpurun_out/bench_c2_f8.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_c2_f8.err
timeout 900 python bench.py --workload c3 --batch 1000000 --steps 2 --warmup 1 > gpurun_out/bench_c3_f8.json 2> gpurun_out/bench_c3_f8.err; echo "c3 rc=$?"; tail -3 gpurun_out/bench_c3_f8.err
python - <<'PY'
import json
for f in ('bench_c2_f8','bench_c3_f8'):
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, 'value', d['value'], 'cpu', d['cpu_baseline'])
PY
