#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_f6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_f6.log; tail -4 gpurun_out/pytest_gpu_f6.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_f6.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_f6.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_default_f6.json 2> gpurun_out/bench_default_f6.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default_f6.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default_f6.json'))
print('c2 value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'full', d['roofline']['full_batch_launches'].get('frac'))
t=d['tall']; print('tall', t['value'], t['ms_per_step'], 'e2e', t['e2e']['value'], 'frac', t['roofline']['frac'])
PY
