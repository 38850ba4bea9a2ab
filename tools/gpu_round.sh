#!/bin/bash
# Round-end GPU cycle on one box: parity suite, default bench line (all five workloads), the CPU arm, ncu launch list of the bench
# command, per-frame ncu counters of every workload (profiles/ncu_counters.json), full captures of the two traversal kernels on C5.
# usage: tools/gpu_round.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
T0=$(date +%s); timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench wall $(( $(date +%s) - T0 )) s"
timeout 300 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --side-steps 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_list_$TAG.log 2>&1
timeout 900 python tools/ncu_counters.py 2>&1 | tail -6
bash tools/gpu_prof.sh $TAG c5 "k_shadow k_trace0"
bash tools/gpu_prof.sh $TAG c2 "k_shadow k_trace0"
