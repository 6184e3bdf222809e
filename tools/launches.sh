#!/bin/bash
# per-kernel durations of a command (ncu launch list, cold-cache and serialised): tools/launches.sh out.csv cmd...
OUT=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT "$@" > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg={}
for r in rows[1:]:
    agg.setdefault(r[ki][:100],[]).append(float(r[vi].replace(",","")))
for k,v in agg.items():
    if k.startswith("void at::") or k.startswith("at::"): continue
    print(f"{sum(v)/len(v)/1000:9.1f} us x{len(v):3d}  {k}")
PY
