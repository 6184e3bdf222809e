"""Aggregate backward at the bench shape (B=32, C=80, N=20736, M=1296, k=9, bf16): fp32-atomic scatter vs the
deterministic 64-bit fixed-point scatter.  usage: python tools/agg_bwd_profile.py [iters]"""
import sys
import torch
sys.path.insert(0, ".")
from gkgnet_b200 import ops
it = int(sys.argv[1]) if len(sys.argv) > 1 else 10
g = torch.Generator(device="cuda").manual_seed(0)
B, G, N, M, D, k = 32, 2, 20736, 1296, 40, 9
C = G * D
x = torch.randn(B, N, C, device="cuda", generator=g).bfloat16().requires_grad_(True)
y = torch.randn(B, M, C, device="cuda", generator=g).bfloat16().requires_grad_(True)
idx = ops.knn_graph(x.detach(), y.detach(), None, groups=G, k=k, dilation=1)
w = torch.randn(B, N, 2 * C, device="cuda", generator=g).bfloat16()
out = ops.mr_aggregate(x, idx, y, groups=G)
for det in (False, True):
    ops.set_deterministic_aggregate(det)
    for _ in range(3):
        out.backward(w, retain_graph=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        out.backward(w, retain_graph=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"deterministic={det}: {e0.elapsed_time(e1) / it * 1e3:.1f} us per backward (autograd call incl. zero-fill / conversion)")
