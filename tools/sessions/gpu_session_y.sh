#!/bin/bash
# C5 shape on one GPU at reduced rows (2M x 256), TRF and dogbox
mkdir -p gpurun_out
for meth in trf dogbox; do
timeout 900 python bench.py --workload c5 --rows 2000000 --method $meth --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c5small_$meth.json 2> gpurun_out/bench_c5small_$meth.err; echo "c5 $meth rc=$?"; tail -3 gpurun_out/bench_c5small_$meth.err
done
python - <<'PY'
import json
for meth in ('trf','dogbox'):
    try:
        d=json.load(open(f'gpurun_out/bench_c5small_{meth}.json'))
        print(meth, 'value', d['value'], d['ms_per_step'], 'frac', d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['launches'], d['config'])
    except Exception as e: print(meth, e)
PY
