#!/bin/bash
# refresh ncu evidence on the final kernels: C3 lin + dogbox round (kbench), C2 launch list
mkdir -p gpurun_out
timeout 300 python tools/kbench.py --workload c3 --B 500000 > gpurun_out/kbench_c3_g4.log 2>&1; cat gpurun_out/kbench_c3_g4.log | cut -c1-400
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lin_kernel -s 2 -c 1 -f -o gpurun_out/prof_lin_c3_r1c python tools/kbench.py --workload c3 --B 500000 --reps 2 > gpurun_out/ncu_lin_c3_g4.log 2>&1; echo "ncu lin c3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:round_kernel -s 2 -c 1 -f -o gpurun_out/prof_round_c3_r1c python tools/kbench.py --workload c3 --B 500000 --reps 2 > gpurun_out/ncu_round_c3_g4.log 2>&1; echo "ncu round c3 rc=$?"
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-tall"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1d.csv $BENCH > gpurun_out/ncu_launch_g4.log 2>&1; echo "ncu launches rc=$?"
