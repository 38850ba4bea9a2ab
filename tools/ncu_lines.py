"""Aggregate an `ncu --page source --csv --print-source sass,cuda` export by CUDA source line."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr = None, None
agg, samples, texts = collections.Counter(), collections.Counter(), {}
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
    if hdr is None or cur_file is None or hdr[0] != "Line No":
        continue
    d = dict(zip(hdr, r))
    try:
        line = int(d["Line No"])
        ie = int(float(d.get("Instructions Executed") or 0))
        sm = int(float(d.get("# Samples") or 0))
    except ValueError:
        continue
    key = (cur_file.split("/")[-1], line)
    agg[key] += ie
    samples[key] += sm
    texts[key] = d["Source"][:100]
tot, ts = sum(agg.values()), sum(samples.values())
print("total warp-inst", tot, "samples", ts)
byfile = collections.Counter()
for k, v in agg.items():
    byfile[k[0]] += v
print({k: f"{100*v/tot:.1f}%" for k, v in byfile.items()})
for k, v in agg.most_common(top):
    print(f"{100*v/tot:5.1f}% inst {100*samples[k]/max(ts,1):5.1f}% smp  {k[0]}:{k[1]:4d}  {texts[k]}")
