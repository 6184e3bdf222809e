#!/bin/bash
# Standard GPU pass (run under gpurun): tests, smoke, bench, ncu launch list + full capture.
# Usage: tools/gpu_check.sh [tag] [extra bench args]
TAG=${1:-r01}
shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu" 
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.txt
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 "$@" 2>&1 | tail -3 | tee gpurun_out/${TAG}_bench.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gkg|knn|mr_aggregate|tc_prepare' -c 120 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu "$@" > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_bench.log
echo "== ncu full (knn + aggregate kernels)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'knn_tc_kernel|knn_finalize|mr_aggregate' -s 12 -c 4 \
    -o gpurun_out/${TAG}_prof -f python bench.py --steps 2 --warmup 3 --no-cpu "$@" > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out
