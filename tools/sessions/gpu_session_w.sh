#!/bin/bash
# 8-GPU run: default bench (C2 weak + tall C4 strong) and C3 (10M problems over 8 GPUs)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/bench_8gpu_w.json 2> gpurun_out/bench_8gpu_w.err; echo "8gpu rc=$?"; tail -3 gpurun_out/bench_8gpu_w.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload c3 --batch 1250000 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c3_8gpu_w.json 2> gpurun_out/bench_c3_8gpu_w.err; echo "c3 8gpu rc=$?"; tail -3 gpurun_out/bench_c3_8gpu_w.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --no-cpu-baseline > gpurun_out/bench_4gpu_w.json 2> gpurun_out/bench_4gpu_w.err; echo "4gpu rc=$?"; tail -3 gpurun_out/bench_4gpu_w.err
python - <<'PY'
import json
for f in ('bench_8gpu_w','bench_c3_8gpu_w','bench_4gpu_w'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'])
        if 'tall' in d: print('  tall', d['tall']['value'], d['tall']['ms_per_step'], d['tall']['e2e']['value'], d['tall']['roofline']['frac'], d['tall']['roofline']['avg_launch_ms'])
    except Exception as e: print(f, e)
PY
