#!/bin/bash
# tall Gauss-Newton shortcut + new fun kernel: GPU tests, smoke, C4 bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_x.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_x.log; tail -3 gpurun_out/pytest_gpu_x.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_x.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_x.log
timeout 900 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/bench_c4_x.json 2> gpurun_out/bench_c4_x.err; echo "c4 rc=$?"; tail -2 gpurun_out/bench_c4_x.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c4_x.json'))
print('c4 value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['launches'], d['config'].get('iterations_per_step'), d['config'].get('nfev_per_step'))
PY
