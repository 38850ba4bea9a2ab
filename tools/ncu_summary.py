"""Summarise an .ncu-rep (one `ncu --set full` capture) as a short markdown block for profiles/.
usage: python tools/ncu_summary.py report.ncu-rep [more.ncu-rep ...] > profiles/xyz.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "block limit (registers)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC (active)"), ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp inst"),
    ("smsp__thread_inst_executed_per_inst_executed.pct", None),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "memory throughput %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_bytes.sum", "L1TEX bytes"), ("l1tex__t_sector_hit_rate.pct", "L1TEX hit %"),
    ("smsp__inst_executed_op_local_ld.sum", "local loads (warp inst)"), ("smsp__inst_executed_op_local_st.sum", "local stores (warp inst)"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp cycles / issued inst"),
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [dict(zip(hdr, zip(units, r))) for r in rows[2:]]


def main():
    for path in sys.argv[1:]:
        for k in raw(path):
            name = k["Kernel Name"][1]
            print(f"### {name}  ({path.split('/')[-1]})\n")
            for key, label in KEYS:
                if key in k and label:
                    u, v = k[key]
                    print(f"- {label}: {v} {u}".rstrip())
            try:  # L2 traffic (SURVEY 8d asks for achieved L2 GB/s next to HBM GB/s): sectors are 32 B
                sectors = float(k["lts__t_sectors.sum"][1])
                unit, dur = k["gpu__time_duration.sum"]
                sec = float(dur) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[unit]
                dr, dw = k["dram__bytes_read.sum"], k["dram__bytes_write.sum"]
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                dram = float(dr[1]) * scale[dr[0]] + float(dw[1]) * scale[dw[0]]
                print(f"- L2 traffic: {sectors * 32 / 1e6:.1f} MB = {sectors * 32 / sec / 1e9:.0f} GB/s;  DRAM traffic: {dram / 1e6:.1f} MB = {dram / sec / 1e9:.0f} GB/s")
            except (KeyError, ValueError):
                pass
            pipes = []
            for key, (u, v) in k.items():  # which execution pipe is the busiest (alu: PRMT/FMNMX/LOP3/SEL, fma: FFMA/FMUL/IMAD)
                if key.startswith("sm__inst_executed_pipe_") and key.endswith(".avg.pct_of_peak_sustained_active"):
                    try:
                        pipes.append((float(v), key[len("sm__inst_executed_pipe_"):-len(".avg.pct_of_peak_sustained_active")]))
                    except ValueError:
                        pass
            pipes.sort(reverse=True)
            if pipes:
                print("- pipe utilisation (% of peak, instructions executed): " + ", ".join(f"{n} {p:.1f}" for p, n in pipes[:6]))
            stalls = []
            for key, (u, v) in k.items():
                if key.startswith("smsp__average_warps_issue_stalled_") and key.endswith("_per_issue_active.ratio") and "not_issued" not in key:
                    try:
                        stalls.append((float(v), key[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            tot = sum(s for s, _ in stalls) or 1.0
            print("- warp stall reasons (warps per issue-active cycle): " + ", ".join(f"{n} {s:.2f} ({100*s/tot:.0f}%)" for s, n in stalls[:8]))
            print()


if __name__ == "__main__":
    main()
