#!/bin/bash
# Standard GPU cycle (run under gpurun): parity tests, bench, ncu launch list, ncu full capture of the top kernels.
# usage: tools/gpu_cycle.sh TAG [tests|notests] [kernel-regex ...]
TAG=${1:-x}; shift
TESTS=${1:-tests}; shift
mkdir -p gpurun_out
if [ "$TESTS" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_$TAG.log
  tail -3 gpurun_out/pytest_$TAG.log
fi
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_$TAG.log 2>&1
for K in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 6 -c 1 -f -o gpurun_out/prof_${K}_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${K}_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_${K}_$TAG.log
done
