#!/usr/bin/env python
"""Per-SASS-instruction stall samples of one kernel in an .ncu-rep.
usage: python tools/ncu_sass.py rep kernel-regex [min_exec_M] [top]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1:3]
minex = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name": break
    if len(r) >= len(hdr) and r[0] != "Address": data.append(r)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in data)
totex = sum(int(r[ix["Instructions Executed"]]) for r in data)
print(f"{len(data)} SASS lines, {tot} samples, {totex/1e6:.1f}M warp instructions")
agg = {}
for h in stalls:
    agg[h[6:]] = sum(int(r[ix[h]]) for r in data)
print("stalls:", ", ".join(f"{k}={v} ({100*v/max(tot,1):.0f}%)" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
out = []
for i, r in enumerate(data):
    s = int(r[ix["# Samples"]]); ex = int(r[ix["Instructions Executed"]])
    if ex / 1e6 < minex: continue
    st = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    out.append((s, f"{i:5d} {s:6d} {ex/1e6:7.2f}M  {r[1].strip()[:78]:78s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}"))
if "--seq" in sys.argv:
    for s, line in out: print(line)
else:
    for s, line in sorted(out, key=lambda t: -t[0])[:top]: print(line)
