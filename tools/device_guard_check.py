#!/usr/bin/env python
"""ADVICE r1 (device guards): a model living on cuda:1 while cuda:0 is the current device must run on cuda:1 and give
the same results as on cuda:0 (needs two GPUs, one process):  python tools/device_guard_check.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gkgnet_b200 as G
from gkgnet_b200 import _lib, ops

assert torch.cuda.device_count() >= 2, "needs two GPUs"
torch.cuda.set_device(0)
torch.manual_seed(0)
x = torch.randn(2, 2304, 80).to(torch.bfloat16)
y = torch.randn(2, 576, 80).to(torch.bfloat16)
res = []
for dev in ("cuda:0", "cuda:1"):
    xd, yd = x.to(dev), y.to(dev)
    idx = ops.knn_graph(xd, yd, None, groups=2, k=9, dilation=1, algo=_lib.KNN_TCGEN05)
    agg = ops.mr_aggregate(xd, idx, yd, groups=2)
    w = torch.randn(160, 40, 1, 1, device=dev) * 0.1
    fc = ops.grouped_fc(agg, ops.grouped_fc_weights(w), torch.zeros(160, device=dev), "gelu")
    z = fc.view(2, 48, 48, 160).permute(0, 3, 1, 2)
    bn = ops.batch_norm_train(z, torch.ones(160, device=dev), torch.zeros(160, device=dev), None, None, 0.1, 1e-5, act="gelu")
    assert idx.device == xd.device and bn.device == xd.device
    assert torch.cuda.current_device() == 0
    res.append((idx.cpu(), agg.cpu(), fc.cpu(), bn.cpu()))
for a, b in zip(*res):
    assert torch.equal(a, b)
print("device_guard_check: kNN (tcgen05), aggregate, grouped FC and batch norm on cuda:1 with cuda:0 current == the same on cuda:0")
