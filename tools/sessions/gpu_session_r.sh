#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python bench.py --no-tall --no-cpu-baseline > gpurun_out/bench_c2_r$i.json 2> gpurun_out/bench_c2_r$i.err; echo "c2 rc=$?"; tail -3 gpurun_out/bench_c2_r$i.err
done
python - <<'PY'
import json
for f in ('bench_c2_r1','bench_c2_r2'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, 'value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], d['clocks'])
    except Exception as e: print(f, e)
PY
