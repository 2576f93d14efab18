#!/bin/bash
# final records for profiles/: default bench (C2 + tall + parity), reference arm, C3 at 10M
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_default_f9.json 2> gpurun_out/bench_default_f9.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default_f9.err
timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_f9.json 2> gpurun_out/bench_ref_f9.err; echo "ref rc=$?"
timeout 900 python bench.py --workload c3 --steps 2 --warmup 1 > gpurun_out/bench_c3_f9.json 2> gpurun_out/bench_c3_f9.err; echo "c3 rc=$?"; tail -3 gpurun_out/bench_c3_f9.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default_f9.json'))
print('c2 value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'full', d['roofline']['full_batch_launches'].get('frac'), 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['parity_vs_gpu'])
t=d['tall']; print('tall', t['value'], t['ms_per_step'], 'e2e', t['e2e']['value'], 'frac', t['roofline']['frac'])
c=json.load(open('gpurun_out/bench_c3_f9.json')); print('c3', c['value'], c['ms_per_step'], 'e2e', c['e2e']['value'], 'frac', c['roofline']['frac'], c['cpu_baseline']['parity_vs_gpu'])
PY
