#!/bin/bash
# final validation of the round on the final code
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_f7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_f7.log; tail -3 gpurun_out/pytest_gpu_f7.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_f7.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_f7.log
timeout 1200 python bench.py > gpurun_out/bench_default_f7.json 2> gpurun_out/bench_default_f7.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default_f7.err
timeout 600 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/bench_c4_f7.json 2> gpurun_out/bench_c4_f7.err; echo "c4 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default_f7.json'))
print('c2 value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'full', d['roofline']['full_batch_launches'].get('frac'), 'cb', d['roofline']['callbacks_ms_per_step'], 'cpu', d['cpu_baseline']['value'])
t=d['tall']; print('tall', t['value'], t['ms_per_step'], 'e2e', t['e2e']['value'], 'frac', t['roofline']['frac'], t['cpu_baseline']['value'])
PY
