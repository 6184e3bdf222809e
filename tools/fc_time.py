#!/usr/bin/env python
"""Grouped-FC kernels at the GKGNet-576 training shapes (16 images): forward, data gradient, weight gradient, with the
HBM roofline of each (algorithmic bytes / measured peak)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gkgnet_b200 import _lib, ops

PEAK = 6525.9e9
dev = torch.device("cuda")
lib = _lib.load()

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3   # us

for name, rows, c2 in (("stage1", 16 * 20736, 160), ("stage2", 16 * 5184, 320), ("stage3", 16 * 1296, 800), ("stage4", 16 * 324, 1280)):
    cg = c2 // 4
    x = torch.randn(rows, c2, device=dev).to(torch.bfloat16)
    go = torch.randn(rows, c2, device=dev).to(torch.bfloat16)
    w = torch.randn(c2, cg, 1, 1, device=dev) * 0.05
    bias = torch.zeros(c2, device=dev)
    w_op = ops.grouped_fc_weights(w)
    w_t = ops.grouped_fc_weights(w, transpose=True)
    gw = torch.zeros(4, cg, cg, device=dev)
    t_f = timeit(lambda: ops.grouped_fc(x, w_op, bias, None))
    t_g = timeit(lambda: ops.grouped_fc(x, w_op, bias, "gelu"))
    t_d = timeit(lambda: ops.grouped_fc(go, w_t, bias, None))
    st = torch.cuda.current_stream().cuda_stream
    t_w = timeit(lambda: (gw.zero_(), lib.gkg_grouped_fc_wgrad(go.data_ptr(), x.data_ptr(), gw.data_ptr(), rows, c2, st)))
    t_p = timeit(lambda: ops.grouped_fc_weights(w))
    by = 2 * rows * c2 * 2
    print(f"{name}: rows={rows} 2C={c2}  fwd {t_f:6.1f} us ({by / (t_f * 1e-6) / PEAK:4.0%})  fwd+gelu {t_g:6.1f} us ({by / (t_g * 1e-6) / PEAK:4.0%})  "
          f"dgrad {t_d:6.1f} us  wgrad {t_w:6.1f} us ({by / (t_w * 1e-6) / PEAK:4.0%})  pack {t_p:5.1f} us  [HBM floor {by / PEAK * 1e6:5.1f} us]")
