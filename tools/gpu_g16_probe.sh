#!/bin/bash
# ncu counters of k_shadow on C5 for library variants: usage tools/gpu_g16_probe.sh "default g16"
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_bytes.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum
for v in ${1:-default}; do
  lib=ray_tracing_gallery_b200/csrc/libb200rt.so; [ $v != default ] && lib=ray_tracing_gallery_b200/csrc/libb200rt_$v.so
  B200RT_LIB=$PWD/$lib timeout 300 ncu --metrics $M --clock-control none -k regex:k_shadow -s 4 -c 1 --csv --log-file gpurun_out/probe_$v.csv \
    python bench.py --workload c5 --only c5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
  python - $v gpurun_out/probe_$v.csv <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[2], errors="replace")) if r]
h = next(i for i, r in enumerate(rows) if "Metric Name" in r)
c = {n: i for i, n in enumerate(rows[h])}
out = {r[c["Metric Name"]]: (r[c["Metric Value"]], r[c["Metric Unit"]]) for r in rows[h + 1:] if len(r) >= len(rows[h])}
print(sys.argv[1], " | ".join(f"{k.split('__')[-1]}={v[0]}{v[1]}" for k, v in out.items()))
PY
done
