#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_g3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_g3.log; tail -3 gpurun_out/pytest_gpu_g3.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_g3.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_g3.log
timeout 600 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/bench_c4_g3.json 2> gpurun_out/bench_c4_g3.err; echo "c4 rc=$?"; tail -2 gpurun_out/bench_c4_g3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c4_g3.json'))
print('c4 value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['config']['iterations_per_step'], d['config']['nfev_per_step'], d['config']['status'])
PY
