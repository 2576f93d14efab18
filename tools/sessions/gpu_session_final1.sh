#!/bin/bash
# round-end style validation: GPU tests, smoke, default bench (c2 + tall), reference arm, benchmark table, C4 launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_f1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_f1.log; tail -3 gpurun_out/pytest_gpu_f1.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_f1.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_f1.log
timeout 1200 python bench.py > gpurun_out/bench_default_f1.json 2> gpurun_out/bench_default_f1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default_f1.err
timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_f1.json 2> gpurun_out/bench_ref_f1.err; echo "ref rc=$?"; tail -3 gpurun_out/bench_ref_f1.err
timeout 900 python bench.py --workload c3 --steps 2 --warmup 1 > gpurun_out/bench_c3_f1.json 2> gpurun_out/bench_c3_f1.err; echo "c3 rc=$?"; tail -3 gpurun_out/bench_c3_f1.err
timeout 600 python benchmarks/run_benchmarks.py gpurun_out/benchmark_table_exact.txt > /dev/null 2> gpurun_out/benchmark_table.err; echo "table rc=$?"; tail -2 gpurun_out/benchmark_table.err; head -30 gpurun_out/benchmark_table_exact.txt
BENCH="python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_c4_r1b.csv $BENCH > gpurun_out/ncu_launch_c4_f1.log 2>&1; echo "ncu c4 launches rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default_f1.json'))
print('c2 value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'full', d['roofline']['full_batch_launches'].get('frac'), 'cpu', d['cpu_baseline'])
t=d['tall']; print('tall', t['value'], t['ms_per_step'], 'e2e', t['e2e']['value'], 'frac', t['roofline']['frac'], t['cpu_baseline'])
r=json.load(open('gpurun_out/bench_ref_f1.json')); print('ref', r['value'], r['tall']['value'])
c=json.load(open('gpurun_out/bench_c3_f1.json')); print('c3', c['value'], c['ms_per_step'], 'e2e', c['e2e']['value'], 'frac', c['roofline']['frac'], 'full', c['roofline']['full_batch_launches'].get('frac'), c['cpu_baseline'])
PY
