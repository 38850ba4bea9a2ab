#!/bin/bash
# One GPU call: the whole parity suite (no -x: every failure is reported), the default bench line, and the two dynamic
# workloads whose frames now overlap their scene updates (double-buffered TLAS).   usage: tools/gpu_verify.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
timeout 540 python -m pytest tests -m gpu -q --timeout 150 2>&1 | tail -40 > gpurun_out/pytest_$TAG.log
tail -5 gpurun_out/pytest_$TAG.log
timeout 240 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json
timeout 120 python bench.py --workload default --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_default_$TAG.json 2> gpurun_out/bench_default_$TAG.err
cat gpurun_out/bench_default_$TAG.json
timeout 150 python bench.py --workload c4 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c4_$TAG.json 2> gpurun_out/bench_c4_$TAG.err
cat gpurun_out/bench_c4_$TAG.json
