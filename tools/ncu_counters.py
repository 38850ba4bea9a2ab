"""Per-frame ncu counters of every frame kernel on every BASELINE workload -> profiles/ncu_counters.json (read by bench.py's
roofline block: instruction counts and L2 / DRAM bytes per frame are properties of workload + binary; bench.py divides
them by the kernel durations it measures live).  Run on the GPU box:

    python tools/ncu_counters.py [c1 c2 c3 c4 c5]        # writes gpurun_out/ncu_counters.json + the raw CSVs

One `ncu --metrics ... --clock-control none` pass per workload over `bench.py --only W --steps 1 --warmup 3 --no-e2e`; the
launch list is cut into frames at `k_sun_dirs` and the LAST plain frame (no counter / timing instantiation) is summed per
kernel family.  Also keeps the launch list (gpu__time_duration) of that frame."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ("gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,lts__t_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
           "l1tex__data_pipe_lsu_wavefronts.sum")
FAMILIES = ["k_sun_dirs", "k_trace", "k_prep", "k_shadow", "k_resolve", "k_tail", "k_mega"]


def family(name):
    for f in FAMILIES:
        if f in name:
            return f
    return None


def to_float(v, unit):
    x = float(v.replace(",", ""))
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "inst": 1.0, "": 1.0}
    return x * scale.get(unit, 1.0)


def parse(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    launches = {}
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        lid = int(r[col["ID"]])
        d = launches.setdefault(lid, {"name": r[col["Kernel Name"]]})
        d[r[col["Metric Name"]]] = to_float(r[col["Metric Value"]], r[col["Metric Unit"]])
    return [launches[k] for k in sorted(launches)]


def last_plain_frame(launches):
    frames, cur = [], None
    for l in launches:
        fam = family(l["name"])
        if fam is None:
            continue
        if fam == "k_sun_dirs":
            cur = []
            frames.append(cur)
        if cur is not None:
            cur.append(l)
    plain = [f for f in frames if not any("<(bool)1" in l["name"] or "ILb1" in l["name"] for l in f) and len(f) >= 4]
    return plain[-1] if plain else None


def main():
    workloads = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out = {}
    for w in workloads:
        log = os.path.join(ROOT, "gpurun_out", f"ncu_counters_{w}.csv")
        cmd = ["ncu", "--metrics", METRICS, "--clock-control", "none", "--csv", "--log-file", log, "-k", "regex:k_(sun_dirs|trace|prep|shadow|resolve|tail|mega)",
               sys.executable, os.path.join(ROOT, "bench.py"), "--workload", w, "--only", w, "--steps", "1", "--warmup", "3", "--no-cpu-baseline", "--no-e2e"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        if r.returncode != 0:
            print(w, "ncu failed", r.stderr[-500:])
            continue
        frame = last_plain_frame(parse(log))
        if not frame:
            print(w, "no plain frame found")
            continue
        fams = {}
        for l in frame:
            f = fams.setdefault(family(l["name"]), {"warp_inst": 0.0, "thread_inst": 0.0, "lts_bytes": 0.0, "dram_bytes": 0.0, "l1_wavefronts": 0.0, "duration_ms": 0.0, "launches": 0})
            f["warp_inst"] += l.get("smsp__inst_executed.sum", 0.0)
            f["thread_inst"] += l.get("smsp__thread_inst_executed.sum", 0.0)
            f["lts_bytes"] += l.get("lts__t_bytes.sum", 0.0)
            f["dram_bytes"] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
            f["l1_wavefronts"] += l.get("l1tex__data_pipe_lsu_wavefronts.sum", 0.0)
            f["duration_ms"] += l.get("gpu__time_duration.sum", 0.0)
            f["launches"] += 1
        for f in fams.values():
            f["source"] = f"tools/ncu_counters.py, one frame of {w} (ncu --metrics {METRICS.split(',')[1]}..., --clock-control none)"
        out[w] = fams
        tot = sum(f["duration_ms"] for f in fams.values())
        print(w, " ".join(f"{k}: {v['duration_ms']:.3f} ms {v['thread_inst'] / max(v['warp_inst'], 1):.1f} lanes" for k, v in fams.items()), f"| frame {tot:.3f} ms under ncu")
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ncu_counters.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
