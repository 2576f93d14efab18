#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_g1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_g1.log; tail -3 gpurun_out/pytest_gpu_g1.log
timeout 600 python bench.py --workload c5 --rows 2000000 --method dogbox --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/bench_c5small_dogbox_g1.json 2> gpurun_out/bench_c5small_dogbox_g1.err; echo "c5 dogbox rc=$?"; tail -3 gpurun_out/bench_c5small_dogbox_g1.err
timeout 600 python bench.py --workload c4 --method dogbox --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c4_dogbox_g1.json 2> gpurun_out/bench_c4_dogbox_g1.err; echo "c4 dogbox rc=$?"; tail -3 gpurun_out/bench_c4_dogbox_g1.err
python - <<'PY'
import json
for f in ('bench_c5small_dogbox_g1','bench_c4_dogbox_g1'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, 'value', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['launches'], d['config']['status'])
    except Exception as e: print(f, e)
PY
