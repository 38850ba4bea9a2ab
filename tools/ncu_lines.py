"""Aggregate an `ncu --page source --csv --print-source sass,cuda` export by CUDA source line:
share of executed warp instructions, share of stall samples, and the line's dominant stall reasons.
usage: ncu -i report.ncu-rep --page source --csv --print-source sass,cuda > src.csv; python tools/ncu_lines.py src.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr = None, None
agg, samples, texts = collections.Counter(), collections.Counter(), {}
threads = collections.Counter()  # thread instructions executed: / warp instructions = active lanes per instruction
stalls = collections.defaultdict(collections.Counter)
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] in ("Line No", "Address"):
        hdr = r
        continue
    if hdr is None or cur_file is None or hdr[0] != "Line No" or not r[0].isdigit():
        continue
    line = int(r[0])
    col = {name: i for i, name in reversed(list(enumerate(hdr)))}  # first occurrence of duplicated names
    try:
        ie = int(float(r[col["Instructions Executed"]] or 0))
        sm = int(float(r[col["# Samples"]] or 0))
        ti = int(float(r[col["Thread Instructions Executed"]] or 0)) if "Thread Instructions Executed" in col else 0
    except ValueError:
        continue
    key = (cur_file.split("/")[-1], line)
    agg[key] += ie
    threads[key] += ti
    samples[key] += sm
    texts[key] = r[1].strip()[:110]
    for name, i in col.items():
        if name.startswith("stall_") and "Not Issued" not in name:
            try:
                stalls[key][name[6:]] += int(float(r[i] or 0))
            except ValueError:
                pass
tot, ts = sum(agg.values()), sum(samples.values())
print("total warp-inst", tot, "samples", ts, "active lanes per warp instruction %.2f" % (sum(threads.values()) / max(tot, 1)))
# idle lane-slots per line: where the SIMD width is lost (32 * warp instructions - thread instructions)
idle = {k: 32 * agg[k] - threads[k] for k in agg}
tidle = sum(idle.values()) or 1
byfile, byfile_s = collections.Counter(), collections.Counter()
for k, v in agg.items():
    byfile[k[0]] += v
    byfile_s[k[0]] += samples[k]
print({k: f"{100*v/tot:.1f}% inst / {100*byfile_s[k]/max(ts,1):.1f}% smp" for k, v in byfile.items()})
allst = collections.Counter()
for k in stalls:
    allst.update(stalls[k])
st = sum(allst.values()) or 1
print("stall samples:", ", ".join(f"{n} {100*c/st:.0f}%" for n, c in allst.most_common(8)))
order = sorted(agg, key=lambda k: -(agg[k] / max(tot, 1) + samples[k] / max(ts, 1)))
for k in order[:top]:
    why = ", ".join(f"{n} {c}" for n, c in stalls[k].most_common(3) if c)
    print(f"{100*agg[k]/tot:5.1f}% inst {100*samples[k]/max(ts,1):5.1f}% smp {threads[k]/max(agg[k],1):5.1f} lanes {100*idle[k]/tidle:5.1f}% idle  {k[0]}:{k[1]:4d}  {texts[k]}   [{why}]")
