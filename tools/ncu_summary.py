#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per-kernel headline metrics and, for one
kernel, the stall-reason totals and hottest SASS lines.  Usage:
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-regex] [top-N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
for r in rows[2:]:
    print("==", r[ix["Kernel Name"]][:100])
    for w in want:
        if w in ix:
            print(f"   {w:70s} {r[ix[w]]:>16s} {units[ix[w]]}")
if kre:
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
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    print(f"-- {kre}: {len(data)} SASS lines, {tot} samples")
    agg = sorted(((sum(int(r[ix[s]]) for r in data), s) for s in stalls), reverse=True)
    print("   " + ", ".join(f"{s}={v} ({100 * v / max(tot, 1):.0f}%)" for v, s in agg[:8]))
    print("   inst executed (warp):", sum(int(r[ix["Instructions Executed"]]) for r in data))
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:topn]:
        st = sorted(((int(r[ix[s]]), s) for s in stalls), reverse=True)[:2]
        print(f"   {r[ix['# Samples']]:>7s} {r[ix['Instructions Executed']]:>10s}  {r[1].strip()[:64]:64s} "
              + " ".join(f"{s[6:]}={v}" for v, s in st))
