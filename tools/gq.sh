#!/bin/bash
# quick GPU check: tests (optional filter) + bench line.  usage: tools/gq.sh tag [pytest -k expr]
TAG=${1:-q}; K=${2:-}
mkdir -p gpurun_out
if [ -n "$K" ]; then KE=(-k "$K"); else KE=(); fi
timeout 900 python -m pytest tests -x -q -m gpu "${KE[@]}" 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('ms_per_step','phase_ms')}, d['roofline']['frac'], d.get('roofline_agg_fwd',{}).get('frac'), d.get('roofline_agg_bwd',{}).get('frac'))"
