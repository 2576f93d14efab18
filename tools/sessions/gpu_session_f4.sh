#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_f4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_f4.log; tail -3 gpurun_out/pytest_gpu_f4.log
timeout 300 python tools/kbench.py --workload c2 > gpurun_out/kbench_f4.log 2>&1; cat gpurun_out/kbench_f4.log | cut -c1-400
for i in 1 2; do
timeout 600 python bench.py --no-tall --no-cpu-baseline --rounds-log gpurun_out/rounds_c2_f4.csv > gpurun_out/bench_c2_f4_$i.json 2> gpurun_out/bench_c2_f4_$i.err; echo "c2 rc=$?"; tail -2 gpurun_out/bench_c2_f4_$i.err
done
python - <<'PY'
import json
for f in ('bench_c2_f4_1','bench_c2_f4_2'):
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, 'value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'cb ms', d['roofline']['callbacks_ms_per_step'], 'kern ms', d['roofline']['kernels_ms_per_step'])
PY
head -10 gpurun_out/rounds_c2_f4.csv
