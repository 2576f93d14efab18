#!/bin/bash
# final validation of the round: GPU tests, smoke, default bench, reference arm, C3 at the full 10M
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_f5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_f5.log; tail -3 gpurun_out/pytest_gpu_f5.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_f5.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_f5.log
timeout 1200 python bench.py > gpurun_out/bench_default_f5.json 2> gpurun_out/bench_default_f5.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default_f5.err
BLSQ_BENCH_H2D_CHUNKS=8 timeout 600 python bench.py --no-tall --no-cpu-baseline > gpurun_out/bench_c2_h8_f5.json 2> gpurun_out/bench_c2_h8_f5.err; echo "h8 rc=$?"
timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_f5.json 2> gpurun_out/bench_ref_f5.err; echo "ref rc=$?"; tail -3 gpurun_out/bench_ref_f5.err
timeout 900 python bench.py --workload c3 --steps 2 --warmup 1 > gpurun_out/bench_c3_f5.json 2> gpurun_out/bench_c3_f5.err; echo "c3 rc=$?"; tail -3 gpurun_out/bench_c3_f5.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default_f5.json'))
print('c2 value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], 'full', d['roofline']['full_batch_launches'].get('frac'), 'cb', d['roofline']['callbacks_ms_per_step'], 'cpu', d['cpu_baseline']['value'])
t=d['tall']; print('tall', t['value'], t['ms_per_step'], 'e2e', t['e2e']['value'], 'frac', t['roofline']['frac'], t['cpu_baseline']['value'])
h=json.load(open('gpurun_out/bench_c2_h8_f5.json')); print('h2d_chunks=8: value', h['value'], 'e2e', h['e2e']['value'], h['e2e']['ms_per_step'])
r=json.load(open('gpurun_out/bench_ref_f5.json')); print('ref', r['value'], r['tall']['value'])
c=json.load(open('gpurun_out/bench_c3_f5.json')); print('c3', c['value'], c['ms_per_step'], 'e2e', c['e2e']['value'], 'frac', c['roofline']['frac'], 'full', c['roofline']['full_batch_launches'].get('frac'), c['cpu_baseline']['value'])
PY
