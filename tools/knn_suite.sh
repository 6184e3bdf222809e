#!/bin/bash
# kNN timing suite (GKGNet-576 stage shapes, bf16): tools/knn_suite.sh [iters]
IT=${1:-5}
for a in "32 80 144 4 2 9 1" "32 160 72 2 2 9 1" "64 400 36 1 2 9 2" "64 400 36 1 2 9 3" "64 640 18 1 2 9 3"; do
  timeout 100 python tools/knn_case.py $a bf16 $IT 2>&1 | tail -3 | tr '\n' ' '; echo
done
