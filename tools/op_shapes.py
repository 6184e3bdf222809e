#!/usr/bin/env python
"""ATen ops of a GKGNet-576 training step grouped by (op, first input shape, dtypes): where do the non-library,
non-native elementwise passes come from?"""
import os, sys, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gkgnet_b200 as G
G.set_norm_type("BN")
dev = torch.device("cuda")
net = G.GKGNet(choice="s", n_classes=80, size=576, drop_path=0.1).to(dev)
head = G.LabelQueryHead(80, 640).to(dev)
net.train(); head.train()
params = [p for p in list(net.parameters()) + list(head.parameters()) if p.requires_grad]
opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.05, fused=True)
img = torch.randn(16, 3, 576, 576, device=dev)
tgt = (torch.rand(16, 80, device=dev) < 0.04).float()
def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = sum(head.forward_train(net(img), tgt).values())
    loss.backward()
    torch.nn.utils.clip_grad_norm_(params, 5.0)
    opt.step(); opt.zero_grad(set_to_none=True)
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    step(); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.name.startswith("aten::") and ev.self_device_time_total > 0:
        key = f"{ev.name:28s} {str(ev.input_shapes[:2]):60s}"
        agg[key][0] += ev.self_device_time_total; agg[key][1] += 1
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f"{t:8.0f} us x{n:3d}  {k}")
