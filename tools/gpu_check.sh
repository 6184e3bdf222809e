#!/bin/bash
# Standard GPU pass (run under gpurun): tests, smoke, bench lines, ncu launch lists + full captures.
# Usage: tools/gpu_check.sh [tag]      -> gpurun_out/<tag>_*
TAG=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest.txt
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/${TAG}_smoke.txt
echo "== bench (default line: GKGNet-576 training step + stage-1 layer microbench)"
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench.json
echo "== bench --impl reference"
timeout 900 python bench.py --impl reference 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_reference.json
echo "== bench --workload infer"
timeout 600 python bench.py --workload infer --no-cpu 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_infer.json
echo "== ncu launch list of the default bench command (our kernels only)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gkg|knn|mr_aggregate|tc_prepare|grouped_fc|pool_keys|label_|multilabel|neighbor|bn_' -c 2500 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_bench.log | cut -c1-300
echo "== ncu full (layer microbench kernels)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'knn_tc_kernel|knn_finalize|mr_aggregate|tc_prepare_rows|grouped_fc' -s 20 -c 8 \
    -o gpurun_out/${TAG}_prof -f python bench.py --workload layer --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_full.log | cut -c1-300
ls -la gpurun_out | tail -20
