#!/usr/bin/env python
"""Attribute the samples / executed instructions of one kernel in an .ncu-rep to CUDA source lines.

ncu's CSV export of the CUDA source view carries no metrics, so this joins the SASS view
(per-instruction samples, by address) with `nvdisasm -g` line info of the object file.

    python tools/ncu_lines.py gpurun_out/x.ncu-rep gkgnet_b200/csrc/knn_tc_inst_9.o \
        _ZN3gkg2tc13knn_tc_kernelILi11ELi9EEEvNS0_8TcParamsE knn_tc [top-N]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, mangled, kre = sys.argv[1:5]
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubins = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubins[0])], capture_output=True, text=True).stdout

# offset -> source line, for the wanted function only
line_of = {}
cur, active = None, False
for ln in dis.splitlines():
    if ln.startswith(".text."):
        active = ln.strip().rstrip(":") == ".text." + mangled
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) >= len(hdr) and r[0] != "Address":
        data.append(r)
base = int(data[0][0], 16)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for r in data:
    off = int(r[0], 16) - base
    key = line_of.get(off, ("?", 0))
    a = agg[key]
    a[0] += int(r[ix["# Samples"]])
    a[1] += int(r[ix["Instructions Executed"]])
    for s in stalls:
        a[2][s[6:]] += int(r[ix[s]])
ts = sum(a[0] for a in agg.values())
ti = sum(a[1] for a in agg.values())
print(f"{len(data)} SASS lines, {ts} samples, {ti / 1e6:.1f}M warp instructions")
texts = {}
for k in agg:
    if k[0] != "?":
        for root in ("gkgnet_b200/csrc",):
            fp = os.path.join(root, k[0])
            if os.path.isfile(fp) and fp not in texts:
                texts[fp] = open(fp).read().splitlines()
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    fp = os.path.join("gkgnet_b200/csrc", k[0])
    txt = texts.get(fp, [""] * (k[1] + 1))[k[1] - 1].strip()[:70] if k[1] else ""
    st = " ".join(f"{n}={v}" for n, v in a[2].most_common(2))
    print(f"{k[0][:18]:18s}:{k[1]:4d} smp {a[0]:6d} ({100 * a[0] / ts:4.1f}%) inst {a[1] / 1e6:7.1f}M ({100 * a[1] / ti:4.1f}%)  {st:34s} | {txt}")
