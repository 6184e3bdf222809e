#!/bin/bash
# ncu --set full capture of ONE launch of a kernel (regex $2) inside the bench step -> gpurun_out/$1.ncu-rep
TAG=${1:-prof}; KRE=${2:-knn_tc_kernel}; SKIP=${3:-4}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c 1 \
    -o gpurun_out/${TAG} -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
