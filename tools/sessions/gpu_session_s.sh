#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/kbench.py --workload c2 > gpurun_out/kbench_s.log 2>&1; BLSQ_TRF_TWO_KERNELS=0 timeout 300 python tools/kbench.py --workload c2 >> gpurun_out/kbench_s.log 2>&1; cat gpurun_out/kbench_s.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_s.log; tail -4 gpurun_out/pytest_gpu_s.log
for v in 1 0; do
BLSQ_TRF_TWO_KERNELS=$v timeout 600 python bench.py --no-tall --no-cpu-baseline --rounds-log gpurun_out/rounds_c2_s$v.csv > gpurun_out/bench_c2_s$v.json 2> gpurun_out/bench_c2_s$v.err; echo "c2 rc=$?"; tail -3 gpurun_out/bench_c2_s$v.err
done
python - <<'PY'
import json
for f in ('bench_c2_s1','bench_c2_s0'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, 'value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], d['roofline']['round_kernel'], d['clocks'])
    except Exception as e: print(f, e)
PY
