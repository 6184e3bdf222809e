#!/bin/bash
# experiments: time (and optionally profile) the wide-group kNN shapes with library variants built by hand
for v in "" _kc32 _nosplit; do
  export GKG_LIB=$PWD/gkgnet_b200/libgkg_b200$v.so
  [ -f $GKG_LIB ] || continue
  echo "=== variant '$v'"
  timeout 100 python tools/knn_case.py 64 400 36 1 2 9 2 bf16 5 2>&1 | grep -v "^For debugging\|^$" | tail -3 | tr '\n' ' '; echo
  timeout 100 python tools/knn_case.py 64 400 36 1 2 9 3 bf16 5 2>&1 | grep -v "^For debugging\|^$" | tail -3 | tr '\n' ' '; echo
  timeout 100 python tools/knn_case.py 32 160 72 2 2 9 1 bf16 5 2>&1 | grep -v "^For debugging\|^$" | tail -3 | tr '\n' ' '; echo
  timeout 100 python tools/knn_case.py 64 640 18 1 2 9 3 bf16 5 2>&1 | grep -v "^For debugging\|^$" | tail -3 | tr '\n' ' '; echo
done
