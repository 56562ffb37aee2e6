"""Opcode histogram of the library's kernels from `cuobjdump -sass` (static counts): which kernels carry FP64 tensor
instructions (DMMA), asynchronous copies (LDGSTS = cp.async), shuffles, barriers.
usage: python tools/sass_hist.py > profiles/r02_sass_histogram.md"""
import collections
import re
import subprocess
import sys

so = "volumetricinterp_b200/libvolinterp_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
kern = None
hist = collections.defaultdict(collections.Counter)
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        kern = re.sub(r"\(.*", "", name)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", line)
    if m and kern:
        op = m.group(1)
        hist[kern][op.split(".")[0]] += 1
        if op.startswith("DMMA"):
            hist[kern][op] += 1
cols = ["DMMA", "DFMA", "DMUL", "DADD", "MUFU", "SHFL", "LDGSTS", "LDS", "STS", "LDG", "STG", "BAR", "LDL", "STL"]
print("# SASS opcode histogram (static instruction counts, `cuobjdump -sass` of libvolinterp_b200.so, sm_100a)\n")
print("DMMA = FP64 tensor instruction (every `mma.sync.m8n8k4/m16n8k4/m16n8k16.f64` lowers to `DMMA.8x8x4`); LDGSTS = `cp.async`; "
      "LDL/STL = local memory (spills, thread-local arrays).\n")
print("| kernel | total | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
for k in sorted(hist, key=lambda k: -sum(v for o, v in hist[k].items() if "." not in o)):
    h = hist[k]
    tot = sum(v for o, v in h.items() if "." not in o)
    print(f"| `{k}` | {tot} | " + " | ".join(str(h.get(c, 0)) for c in cols) + " |")
kinds = collections.Counter()
for k in hist:
    for o, v in hist[k].items():
        if o.startswith("DMMA."):
            kinds[o] += v
print("\nDMMA forms in the library:", dict(kinds))
