#!/bin/bash
# One gpurun call = one invocation of this script: `gpurun -- tools/gpu_session.sh <tag> <step>...`
# Steps write their logs to gpurun_out/<tag>_*.  Steps:
#   tests        python -m pytest tests -m gpu
#   smoke        __graft_entry__.smoke()
#   bench[=args] python bench.py <args>
#   launches=wl  ncu launch list of `bench.py --workload wl --steps 1 --warmup 1`
#   py=<file>    python <file>
set -u
cd "$(dirname "$0")/.."
TAG=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${TAG}_gpu.txt 2>&1
for step in "$@"; do
  name=${step%%=*}; arg=""; [[ "$step" == *=* ]] && arg=${step#*=}
  case $name in
    tests) python -m pytest tests -m gpu -x -q ${arg:+-k "$arg"} > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log; tail -15 gpurun_out/${TAG}_pytest.log;;
    smoke) python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/${TAG}_smoke.log;;
    bench) k=$(echo "$arg" | tr -c 'a-zA-Z0-9' '_'); python bench.py $arg > gpurun_out/${TAG}_bench_${k}.json 2> gpurun_out/${TAG}_bench_${k}.err; echo "bench $arg rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench_${k}.json; tail -5 gpurun_out/${TAG}_bench_${k}.err;;
    tbench) # tbench=<N>:<bench args>   (one rank per GPU over NCCL); TBENCH_ENV="K=V ..." exported first
       IFS=: read -r ng bargs <<< "$arg"; k=$(echo "$ng $bargs ${TBENCH_TAG:-}" | tr -c 'a-zA-Z0-9' '_')
       python -m torch.distributed.run --nnodes=1 --nproc-per-node $ng --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $ng $bargs > gpurun_out/${TAG}_tbench_${k}.json 2> gpurun_out/${TAG}_tbench_${k}.err; echo "tbench $ng $bargs rc=$?"; tail -c 600 gpurun_out/${TAG}_tbench_${k}.json; tail -3 gpurun_out/${TAG}_tbench_${k}.err;;
    launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches_${arg}.csv python bench.py --workload $arg --steps 1 --warmup 1 --no-cpu-baseline --no-tall > gpurun_out/${TAG}_launches_${arg}.log 2>&1; echo "launches rc=$?";;
    ncu) # ncu=<name>:<kernel regex>:<skip>:<count>:<python args>
       IFS=: read -r nm rx skip cnt pyargs <<< "$arg"
       ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip $skip --launch-count $cnt -o gpurun_out/${TAG}_${nm} -f python $pyargs > gpurun_out/${TAG}_ncu_${nm}.log 2>&1; echo "ncu $nm rc=$?"; tail -3 gpurun_out/${TAG}_ncu_${nm}.log;;
    kbench) for l in $arg; do python tools/kbench.py --lib $l --workload c2; python tools/kbench.py --lib $l --workload c3; done > gpurun_out/${TAG}_kbench.log 2>&1; echo "kbench rc=$?"; cat gpurun_out/${TAG}_kbench.log;;
    py) python $arg > gpurun_out/${TAG}_$(basename $arg .py).log 2>&1; echo "py $arg rc=$?"; tail -30 gpurun_out/${TAG}_$(basename $arg .py).log;;
    sh) bash -c "$arg" > gpurun_out/${TAG}_sh.log 2>&1; echo "sh rc=$?"; tail -30 gpurun_out/${TAG}_sh.log;;
  esac
done
