#!/bin/bash
# ncu full captures (with source) of the named kernels on one workload.  usage: tools/gpu_prof.sh TAG WORKLOAD "k_shadow k_trace0" [extra bench args]
TAG=${1:-x}; W=${2:-c2}; KS=${3:-"k_shadow k_trace0"}; shift 3
mkdir -p gpurun_out
for K in $KS; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 4 -c 1 -f -o gpurun_out/prof_${K}_${W}_$TAG \
    python bench.py --workload $W --only $W --steps 2 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ncu_${K}_${W}_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_${K}_${W}_$TAG.log
done
ls -la gpurun_out/*_$TAG.ncu-rep
