#!/bin/bash
# C5 (BASELINE configs[4]): m = 1e8, n = 256, half the bounds active, 8 GPUs, TRF then dogbox
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --workload c5 --method trf --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c5_trf_8gpu.json 2> gpurun_out/bench_c5_trf_8gpu.err; echo "c5 trf rc=$?"; tail -3 gpurun_out/bench_c5_trf_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --workload c5 --method dogbox --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/bench_c5_dogbox_8gpu.json 2> gpurun_out/bench_c5_dogbox_8gpu.err; echo "c5 dogbox rc=$?"; tail -3 gpurun_out/bench_c5_dogbox_8gpu.err
python - <<'PY'
import json
for meth in ('trf','dogbox'):
    try:
        d=json.loads(open(f'gpurun_out/bench_c5_{meth}_8gpu.json').read().strip().splitlines()[-1])
        print(meth, 'value', d['value'], d['ms_per_step'], 'frac', d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['launches'], d['config'])
    except Exception as e: print(meth, e)
PY
