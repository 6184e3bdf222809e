"""One kNN graph call at a GKGNet-576 stage shape (for ncu captures / timing of the prepare and select kernels).
usage: python tools/knn_one.py stage3|stage4|stage2|stage1 [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gkgnet_b200 import ops  # noqa: E402

SHAPES = {"stage1": (20736, 1296, 80, 1), "stage2": (5184, 1296, 160, 1), "stage3": (1296, None, 400, 2),
          "stage4": (324, None, 640, 3)}
name = sys.argv[1] if len(sys.argv) > 1 else "stage3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
N, M, C, d = SHAPES[name]
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, N, C, device="cuda", generator=g).bfloat16()
y = None if M is None else torch.randn(B, M, C, device="cuda", generator=g).bfloat16()
for _ in range(3):
    ops.knn_graph(x, y, None, groups=2, k=9, dilation=d)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.knn_graph(x, y, None, groups=2, k=9, dilation=d)
e1.record()
torch.cuda.synchronize()
print(f"{name} B={B}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per kNN graph call (prepare + select)")
