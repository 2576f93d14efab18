#!/bin/bash
# One gpurun call: C3 (batched dogbox, 2-point Jacobian) bench + reference arm + bytes traffic of lin kernel.
set -x
mkdir -p gpurun_out
python bench.py --workload c3 --steps 2 --warmup 1 > gpurun_out/bench_c3_b.json 2> gpurun_out/bench_c3_b.err
echo "c3 rc=$?"
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_c2_ref_b.json 2> gpurun_out/bench_c2_ref_b.err
echo "ref rc=$?"
cat gpurun_out/bench_c3_b.json gpurun_out/bench_c2_ref_b.json
tail -5 gpurun_out/bench_c3_b.err
