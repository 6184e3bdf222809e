#!/usr/bin/env python
"""Debug: per-kNN-call counters of the tcgen05 path inside a whole GKGNet-576 forward (rows sent to the
brute-force fix-up / exact re-rank), to check that every layer shape stays on the fast path, and the number of
rows whose neighbour ids differ from the CUDA-core exact fp32 kernel on the same (real-model) features."""
import ctypes, os, struct, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gkgnet_b200 as G
from gkgnet_b200 import _lib, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
size = int(sys.argv[2]) if len(sys.argv) > 2 else 576
lib = _lib.load()
lib.gkg_debug_knn_tc.argtypes = [ctypes.c_int, ctypes.c_void_p]
lib.gkg_debug_knn_tc_stats.argtypes = [ctypes.c_void_p]
orig = ops.knn_graph

def wrapped(x, y=None, relative_pos=None, **kw):
    idx = orig(x, y, relative_pos, **kw)
    torch.cuda.synchronize()
    arr = (ctypes.c_uint * 3)()
    lib.gkg_debug_knn_tc_stats(arr)
    Bx, N, C = x.shape
    M = N if y is None else y.shape[1]
    g = kw.get("groups", 1)
    rows = Bx * g * N
    kw2 = dict(kw); kw2.pop("separable", None); kw2["algo"] = _lib.KNN_EXACT_FP32
    ref = orig(x, y, relative_pos, **kw2)                      # CUDA-core exact kernel on the same features
    differ = int((idx != ref).any(-1).sum())
    print(f"N={N:6d} M={M:6d} D={C // g:4d} k={kw.get('k')} d={kw.get('dilation')} bias={relative_pos is not None} "
          f"rows={rows:8d} fixups={arr[0]:7d} ({100.0 * arr[0] / rows:6.3f}%) reranked={arr[1]:7d} "
          f"max_err={struct.unpack('f', struct.pack('I', arr[2]))[0]:.2e} rows_differ_from_exact_kernel={differ}", flush=True)
    return idx

ops.knn_graph = wrapped
import gkgnet_b200.graph as graph_mod
if hasattr(graph_mod, "ops"):
    graph_mod.ops.knn_graph = wrapped
G.set_norm_type("BN")
dev = torch.device("cuda")
torch.manual_seed(0)
net = G.GKGNet(choice="s", n_classes=80, size=size, drop_path=0.0).to(dev).eval()
img = torch.randn(B, 3, size, size, device=dev)
lib.gkg_debug_knn_tc(-1, None)
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    net(img)
lib.gkg_debug_knn_tc(0, None)
