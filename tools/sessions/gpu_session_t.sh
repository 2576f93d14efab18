#!/bin/bash
# validation + ncu evidence for the reworked batched kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_t.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_t.log; tail -3 gpurun_out/pytest_gpu_t.log
timeout 600 python bench.py --no-tall --no-cpu-baseline > gpurun_out/bench_c2_t.json 2> gpurun_out/bench_c2_t.err; echo "c2 rc=$?"; tail -3 gpurun_out/bench_c2_t.err
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-tall"
$BENCH > gpurun_out/plain_t.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r1c.csv $BENCH > gpurun_out/ncu_launch_t.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lin_kernel -s 2 -c 2 -f -o gpurun_out/prof_lin_r1c $BENCH > gpurun_out/ncu_lin_t.log 2>&1; echo "ncu lin rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:round_kernel -s 4 -c 4 -f -o gpurun_out/prof_round_r1c $BENCH > gpurun_out/ncu_round_t.log 2>&1; echo "ncu round rc=$?"
python - <<'PY'
import json
for f in ('bench_c2_t',):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, 'value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], d['roofline']['full_batch_launches'], d['roofline']['round_kernel'])
    except Exception as e: print(f, e)
PY
