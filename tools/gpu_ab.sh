#!/bin/bash
# A/B of library variants on the GPU box (csrc/Makefile `variant`), one bench line per (variant, workload), then the
# parity suite on the default library.   usage: tools/gpu_ab.sh TAG "variant ..." "workload ..." [tests]
# variant "default" = csrc/libb200rt.so, otherwise csrc/libb200rt_<variant>.so
TAG=${1:-ab}; VARIANTS=${2:-default}; WORKLOADS=${3:-c2}; TESTS=${4:-tests}
mkdir -p gpurun_out
CSRC=ray_tracing_gallery_b200/csrc
for w in $WORKLOADS; do
  for v in $VARIANTS; do
    lib=$CSRC/libb200rt.so; [ "$v" != default ] && lib=$CSRC/libb200rt_$v.so
    steps=100; [ "$w" = c4 ] && steps=20; [ "$w" = c5 ] && steps=5
    B200RT_LIB=$PWD/$lib timeout 200 python bench.py --workload $w --only $w --steps $steps --warmup 5 --no-cpu-baseline \
      > gpurun_out/ab_${TAG}_${v}_$w.json 2> gpurun_out/ab_${TAG}_${v}_$w.err
    python - "$v" "$w" gpurun_out/ab_${TAG}_${v}_$w.json <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
    k = j["roofline"]["kernel_ms_per_frame"]
    print(f"{sys.argv[1]:>20} {sys.argv[2]:>8}: value {j['value']:8.0f} Mrays/s {j['ms_per_step']:.4f} ms | e2e {j['e2e']['value']:8.0f} | trace {k['trace']:.4f} prep {k['prep']:.4f} shadow {k['shadow']:.4f} tail {k['tail']:.4f}")
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e)
PY
  done
done | tee gpurun_out/ab_$TAG.txt
if [ "$TESTS" = tests ]; then
  timeout 400 python -m pytest tests -m gpu -q --timeout 150 2>&1 | tail -25 > gpurun_out/pytest_$TAG.log
  tail -4 gpurun_out/pytest_$TAG.log
fi
