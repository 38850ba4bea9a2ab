#!/bin/bash
# A/B of run-time knobs (environment variables read by rt_create) on the default library.
# usage: tools/gpu_env_ab.sh TAG "name=ENV=VAL ..." "workload ..."     (name "base" with empty assignment = no knob)
TAG=${1:-env}; KNOBS=${2:-base=}; WORKLOADS=${3:-c2}
mkdir -p gpurun_out
for w in $WORKLOADS; do
  for k in $KNOBS; do
    name=${k%%=*}; assign=${k#*=}
    steps=100; [ "$w" = c4 ] && steps=20; [ "$w" = c5 ] && steps=5
    env $assign timeout 200 python bench.py --workload $w --steps $steps --warmup 5 --no-cpu-baseline \
      > gpurun_out/env_${TAG}_${name}_$w.json 2> gpurun_out/env_${TAG}_${name}_$w.err
    python - "$name" "$w" gpurun_out/env_${TAG}_${name}_$w.json <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
    k = j["roofline"]["kernel_ms_per_frame"]
    print(f"{sys.argv[1]:>20} {sys.argv[2]:>8}: value {j['value']:8.0f} Mrays/s {j['ms_per_step']:.4f} ms | e2e {j['e2e']['value']:8.0f} | trace {k['trace']:.4f} prep {k['prep']:.4f} shadow {k['shadow']:.4f} resolve {k['resolve']:.4f} tail {k['tail']:.4f}")
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e)
PY
  done
done | tee gpurun_out/env_$TAG.txt
