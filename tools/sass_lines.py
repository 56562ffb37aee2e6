"""Join an ncu SASS-level source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm -g line info of the same
kernel, and aggregate stall samples / executed instructions per CUDA source line.

    python tools/sass_lines.py <source.csv> <annotated.sass> <mangled-kernel-substring> [top]
"""
import csv
import re
import sys
from collections import defaultdict

src_csv, sass, key = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# 1. offsets -> (file, line) from nvdisasm -g
loc = {}
cur = None
infn = False
for ln in open(sass):
    if ln.startswith(".text.") and ln.rstrip().endswith(":"):
        infn = key in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        loc[int(m.group(1), 16)] = cur
# 2. ncu rows
rows = list(csv.reader(open(src_csv)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
iS, iI = H.index("# Samples"), H.index("Instructions Executed")
base = None
agg = defaultdict(lambda: [0.0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= iI or not r[0].startswith("0x"):
        continue
    a = int(r[0], 16)
    if base is None:
        base = a
    l = loc.get(a - base, ("?", 0))
    agg[l][0] += float(r[iS] or 0)
    agg[l][1] += float(r[iI] or 0)
ts = sum(v[0] for v in agg.values()) or 1.0
ti = sum(v[1] for v in agg.values()) or 1.0
print(f"total samples {ts:.0f}, warp instructions {ti:.0f}")
byfile = defaultdict(lambda: [0.0, 0.0])
for (f, l), v in agg.items():
    byfile[f][0] += v[0]; byfile[f][1] += v[1]
for f, v in sorted(byfile.items(), key=lambda x: -x[1][0]):
    print(f"  {f:28s} samples {100 * v[0] / ts:5.1f} %  instructions {100 * v[1] / ti:5.1f} %")
srccache = {}
def text(f, l):
    import glob, os
    if f not in srccache:
        c = glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "volumetricinterp_b200", "csrc", f))
        srccache[f] = open(c[0]).read().splitlines() if c else []
    s = srccache[f]
    return s[l - 1].strip()[:90] if 0 < l <= len(s) else ""
for (f, l), v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{f}:{l:4d}  samples {100 * v[0] / ts:5.1f} %  instr {100 * v[1] / ti:5.1f} %   {text(f, l)}")
