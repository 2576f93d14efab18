#!/bin/bash
# 2-GPU run: full GPU tests, default bench (C2 + tall C4) over NCCL, C3 split over 2 GPUs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_u.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_u.log; tail -3 gpurun_out/pytest_gpu_u.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_2gpu_u.json 2> gpurun_out/bench_2gpu_u.err; echo "2gpu rc=$?"; tail -3 gpurun_out/bench_2gpu_u.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload c3 --batch 5000000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c3_2gpu_u.json 2> gpurun_out/bench_c3_2gpu_u.err; echo "c3 2gpu rc=$?"; tail -3 gpurun_out/bench_c3_2gpu_u.err
python - <<'PY'
import json
for f in ('bench_2gpu_u','bench_c3_2gpu_u'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'])
        if 'tall' in d: print('  tall', d['tall']['value'], d['tall']['ms_per_step'], d['tall']['e2e']['value'], d['tall']['roofline']['frac'])
    except Exception as e: print(f, e)
PY
