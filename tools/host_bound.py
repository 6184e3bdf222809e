#!/usr/bin/env python
"""Is the GKGNet-576 training step launch-bound?  Host time to enqueue a step vs device time of the step.
usage: python tools/host_bound.py [batch]"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gkgnet_b200 as G

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
G.set_norm_type("BN")
dev = torch.device("cuda")
net = G.GKGNet(choice="s", n_classes=80, size=576, drop_path=0.1).to(dev)
head = G.LabelQueryHead(80, 640).to(dev)
net.train(); head.train()
params = [p for p in list(net.parameters()) + list(head.parameters()) if p.requires_grad]
opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.05, fused=True)
img = torch.randn(B, 3, 576, 576, device=dev)
tgt = (torch.rand(B, 80, device=dev) < 0.04).float()

def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = sum(head.forward_train(net(img), tgt).values())
    loss.backward()
    torch.nn.utils.clip_grad_norm_(params, 5.0)
    opt.step(); opt.zero_grad(set_to_none=True)

for _ in range(4):
    step()
torch.cuda.synchronize()
host, devt = [], []
for _ in range(8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    step()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    host.append((t1 - t0) * 1e3); devt.append(e0.elapsed_time(e1))
print(f"batch {B}: host enqueue {sorted(host)[len(host)//2]:.1f} ms, device span {sorted(devt)[len(devt)//2]:.1f} ms "
      f"-> {B / sorted(devt)[len(devt)//2] * 1e3:.0f} img/s")
