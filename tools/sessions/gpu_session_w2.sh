#!/bin/bash
# 8-GPU run with the final code: default bench (C2 weak + tall C4 strong) and C3
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/bench_8gpu_w2.json 2> gpurun_out/bench_8gpu_w2.err; echo "8gpu rc=$?"; tail -2 gpurun_out/bench_8gpu_w2.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --workload c3 --batch 1250000 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c3_8gpu_w2.json 2> gpurun_out/bench_c3_8gpu_w2.err; echo "c3 8gpu rc=$?"; tail -2 gpurun_out/bench_c3_8gpu_w2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 4 --no-cpu-baseline > gpurun_out/bench_4gpu_w2.json 2> gpurun_out/bench_4gpu_w2.err; echo "4gpu rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_2gpu_w2.json 2> gpurun_out/bench_2gpu_w2.err; echo "2gpu rc=$?"
python - <<'PY'
import json
for f in ('bench_8gpu_w2','bench_c3_8gpu_w2','bench_4gpu_w2','bench_2gpu_w2'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'])
        if 'tall' in d: print('  tall', d['tall']['value'], d['tall']['ms_per_step'], d['tall']['e2e']['value'], d['tall']['roofline']['frac'], d['tall']['roofline']['avg_launch_ms'])
    except Exception as e: print(f, e)
PY
