"""Time gkg_mr_aggregate_fwd (bf16) at the GKGNet-576 stage shapes.  Run under gpurun; set
GKG_AGG_OLD=1 to time the previous (CTA-tiled) shared-memory kernel for an A/B comparison.

usage: python tools/agg_time.py [--train] [--only=I]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gkgnet_b200 import ops  # noqa: E402

SHAPES = [  # name, B, G, N, M, D, k
    ("stage1 r=4 (bench)", 32, 2, 20736, 1296, 40, 9),
    ("stage2 r=2", 32, 2, 5184, 1296, 80, 9),
    ("stage3 r=1 self", 32, 2, 1296, 1296, 200, 9),
    ("stage1 k=18", 32, 2, 20736, 1296, 40, 18),
    ("stage1 448", 32, 2, 12544, 784, 40, 9),
]


def main():
    train = "--train" in sys.argv
    g = torch.Generator(device="cuda").manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    print("path:", "old CTA-tiled" if os.environ.get("GKG_AGG_OLD") else "warp-autonomous", "| argmax:", train)
    only = [int(a.split("=")[1]) for a in sys.argv if a.startswith("--only=")]
    for i, (name, B, G, N, M, D, k) in enumerate(SHAPES):
        if only and i not in only:
            continue
        C = G * D
        x = torch.randn(B, N, C, device="cuda", generator=g).bfloat16().requires_grad_(train)
        self_keys = "self" in name
        y = None if self_keys else torch.randn(B, M, C, device="cuda", generator=g).bfloat16()
        idx = torch.randint(0, M, (B * G, N, k), device="cuda", generator=g, dtype=torch.int32)
        for _ in range(3):
            ops.mr_aggregate(x, idx, y, groups=G)
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.mr_aggregate(x, idx, y, groups=G)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        med = ts[len(ts) // 2]
        by = 2 * B * C * N + (0 if self_keys else 2 * B * C * M) + 4 * B * G * N * k + 2 * B * 2 * C * N + (B * C * N if train else 0)
        print(f"{name:22s} B={B} N={N} M={M} D={D} k={k}: {med * 1e3:7.1f} us (min {ts[0] * 1e3:.1f})  "
              f"{by / med / 1e6:7.0f} GB/s algorithmic")


if __name__ == "__main__":
    main()
