#!/bin/bash
# A/B of environment knobs on one box: usage tools/gpu_env_ab2.sh TAG "name:VAR=val,VAR2=val ..." "workloads"
TAG=${1:-env}; VARIANTS=${2:-"default:"}; WORKLOADS=${3:-c2}
mkdir -p gpurun_out
for w in $WORKLOADS; do
  for v in $VARIANTS; do
    name=${v%%:*}; envs=$(echo "${v#*:}" | tr ',' ' ')
    steps=100; [ "$w" = c4 ] && steps=20; [ "$w" = c5 ] && steps=5
    env $envs timeout 300 python bench.py --workload $w --only $w --steps $steps --warmup 5 --no-cpu-baseline > gpurun_out/ab_${TAG}_${name}_$w.json 2> gpurun_out/ab_${TAG}_${name}_$w.err
    python - "$name" "$w" gpurun_out/ab_${TAG}_${name}_$w.json <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
    k = j["roofline"]["kernel_ms_per_frame"]; pr = j["roofline"]["per_ray"]
    print(f"{sys.argv[1]:>14} {sys.argv[2]:>4}: value {j['value']:8.0f} Mrays/s {j['ms_per_step']:.4f} ms | trace {k['trace']:.4f} prep {k['prep']:.4f} shadow {k['shadow']:.4f} tail {k['tail']:.4f} | nodes/ray {pr['nodes']:.2f} inst {pr['instances']:.2f} tris {pr['triangles']:.2f}")
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e)
PY
  done
done | tee gpurun_out/ab_$TAG.txt
