#!/bin/bash
# Round-end GPU cycle: parity suite, default bench line, tail-policy A/B on C3, ncu launch list, ncu full captures of the
# two traversal kernels.   usage: tools/gpu_final.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q --timeout 150 2>&1 | tail -30 > gpurun_out/pytest_$TAG.log
tail -4 gpurun_out/pytest_$TAG.log
timeout 200 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json
bash tools/gpu_env_ab.sh $TAG "coop=B200RT_SPLIT_TAIL=0 adaptive=X=0" "c3"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_$TAG.log 2>&1
for K in k_shadow k_trace0; do
  timeout 240 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 6 -c 1 -f -o gpurun_out/prof_${K}_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${K}_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_${K}_$TAG.log
done
ls -la gpurun_out/*.ncu-rep
