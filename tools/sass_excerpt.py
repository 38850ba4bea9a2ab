"""SASS evidence for profiles/: the node visit (eight LDG.E.128.CONSTANT + the 8-child slab test on packed bf16) and the leaf loop of a traversal
kernel, cut out of `cuobjdump -sass libb200rt.so`, plus an opcode histogram of the whole kernel.
usage: python tools/sass_excerpt.py k_shadow > profiles/r03_k_shadow_sass.md"""
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
      "not a contraction (north_star).  The packed-bf16 instructions of the box test are HFMA2.BF16_V2, VHMNMX.BF16_V2 (3-input min / max), "
      "HMNMX2.BF16_V2, HADD2.BF16_V2 and F2FP.BF16.F32.PACK_AB.\n")
# node visit = from the node's first load (the first place with several node loads close together that is followed by the packed
# bf16 arithmetic) to the instruction that combines the hit mask with the node's internal-child mask
idx = [i for i, l in enumerate(inst) if "LDG.E.128.CONSTANT" in l]
start = None
for a in idx:
    window = inst[a:a + 120]
    if sum("LDG.E.128.CONSTANT" in w for w in window) >= 8 and any("HFMA2.BF16" in w for w in window):
        start = a
        break
if start is not None:
    end = start
    seen_h = 0
    for j in range(start, min(start + 420, len(inst))):
        if "HADD2.BF16" in inst[j] or ("HFMA2.BF16" in inst[j] and ", 1, 1, -" in inst[j]):
            seen_h += 1
        end = j
        if seen_h >= 4 and ("BRA" in inst[j] or "BSYNC" in inst[j]):
            break
    body = inst[start:end + 1]
    c = collections.Counter(re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", l).group(1) for l in body)
    print(f"## Node visit: {len(body)} instructions from the first of the eight `LDG.E.128.CONSTANT` (ld.global.nc.v4: the node's first 128-byte line — header, "
          "then the near and far bf16 planes of the ray's octant) to the hit mask of the 8-child slab test and its split into internal / leaf children\n")
    print("Mix: " + ", ".join(f"{k} {v}" for k, v in c.most_common(16)) + "\n")
    print("```")
    for l in body:
        print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", l).rstrip())
    print("```")
