"""Run one kNN shape through the tcgen05 path: compare with the exact kernel, time prepare / select.
usage: python tools/knn_case.py B C side r G k d [dtype] [iters] [skip] [ga] [flags]
(flags = -2: the selection epilogue drains the accumulators without ranking them -- timing experiments only)"""
import sys
import time

import torch

sys.path.insert(0, ".")
from gkgnet_b200 import _lib, ops
from gkgnet_b200.pos_embed import relative_pos_table

B, C, side, r, G, k, d = [int(a) for a in sys.argv[1:8]]
dtype = {"bf16": torch.bfloat16, "f32": torch.float32}[sys.argv[8] if len(sys.argv) > 8 else "bf16"]
iters = int(sys.argv[9]) if len(sys.argv) > 9 else 5
dbg = {}
if len(sys.argv) > 10:
    dbg["skip"] = int(sys.argv[10])
if len(sys.argv) > 11:
    dbg["ga"] = int(sys.argv[11])
if len(sys.argv) > 12:
    dbg["flags"] = int(sys.argv[12])
n = side * side
torch.manual_seed(0)
x4 = torch.randn(B, C, side, side, device="cuda")
x = x4.permute(0, 2, 3, 1).reshape(B, n, C).to(dtype).contiguous()
y = None
if r > 1:
    y = torch.nn.functional.avg_pool2d(x4, r, r).permute(0, 2, 3, 1).reshape(B, -1, C).to(dtype).contiguous()
rel = relative_pos_table(C, n, r).cuda()
sep = ops.fit_separable_bias(rel)
print("sep", None if sep is None else sep[2:], "N", n, "M", n // (r * r), "D", C // G, flush=True)
info = dict(dbg)
a = ops.knn_graph(x, y, rel, groups=G, k=k, dilation=d, algo=_lib.KNN_TCGEN05, separable=sep, debug=info)
torch.cuda.synchronize()
print("stats", info["stats"], flush=True)
b = ops.knn_graph(x, y, rel, groups=G, k=k, dilation=d, algo=_lib.KNN_EXACT_FP32)
torch.cuda.synchronize()
diff = (a != b).any(-1).sum().item()
print("rows differing from the exact kernel:", diff, "of", a.shape[0] * a.shape[1])
if iters > 0:
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for it in range(2):
        ops.knn_graph(x, y, rel, groups=G, k=k, dilation=d, algo=_lib.KNN_TCGEN05, separable=sep,
                      debug=dict(dbg) if dbg else None)
    torch.cuda.synchronize()
    ev[0].record()
    for it in range(iters):
        ops.knn_graph(x, y, rel, groups=G, k=k, dilation=d, algo=_lib.KNN_TCGEN05, separable=sep,
                      debug=None)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / iters
    M = n // (r * r)
    print(f"knn_graph (prepare + select): {ms:.3f} ms  -> {2 * B * n * M * C / ms / 1e9:.1f} TF/s algorithmic")
