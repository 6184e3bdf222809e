#!/usr/bin/env python
"""Debug: run the tcgen05 kNN on the bench shape (smaller batch) and print the kernel's counters
(rows sent to the exact fix-up kernel, rows re-ranked exactly, max |approx - exact|)."""
import ctypes, struct, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gkgnet_b200 import _lib, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
lib = _lib.load()
x, y, rel = bench.make_inputs(B, "cuda", torch.bfloat16, 0)
x, y, rel = x.cuda(), y.cuda(), rel.cuda()
sep = ops.fit_separable_bias(rel)
lib.gkg_debug_knn_tc.argtypes = [ctypes.c_int, ctypes.c_void_p]
lib.gkg_debug_knn_tc(-1, None)
idx = ops.knn_graph(x, y, rel, groups=2, k=9, dilation=1, separable=sep)
torch.cuda.synchronize()
arr = (ctypes.c_uint * 3)()
lib.gkg_debug_knn_tc_stats.argtypes = [ctypes.c_void_p]
lib.gkg_debug_knn_tc_stats(arr)
rows = B * 2 * x.shape[1]
print({"rows": rows, "fixups": arr[0], "reranked": arr[1], "max_err": struct.unpack("f", struct.pack("I", arr[2]))[0]})
lib.gkg_debug_knn_tc(0, None)
ref = ops.knn_graph(x, y, rel, groups=2, k=9, dilation=1, algo=_lib.KNN_EXACT_FP32)
print("equal to exact kernel:", bool(torch.equal(idx, ref)), "rows differ:", int((idx != ref).any(-1).sum()))
for _ in range(3):
    ops.knn_graph(x, y, rel, groups=2, k=9, dilation=1, separable=sep)
torch.cuda.synchronize()
t0 = time.time()
for _ in range(5):
    ops.knn_graph(x, y, rel, groups=2, k=9, dilation=1, separable=sep)
torch.cuda.synchronize()
print("knn_graph ms (incl. prepare, alloc):", (time.time() - t0) / 5 * 1e3)
