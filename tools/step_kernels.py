#!/usr/bin/env python
"""Warm per-kernel device time of one GKGNet-576 training step (torch.profiler / CUPTI, eager launch;
the kernels are the ones the captured step replays).  usage: python tools/step_kernels.py [batch] [rows]"""
import os, re, sys, collections
import torch
from torch.profiler import profile, ProfilerActivity
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gkgnet_b200 as G

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
TOP = int(sys.argv[2]) if len(sys.argv) > 2 else 45
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
STEPS = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(STEPS):
        step()
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        t = tot[ev.name[:110]]
        t[0] += 1; t[1] += ev.device_time
s = sum(v[1] for v in tot.values())
print(f"# batch {B}: {s / STEPS / 1e3:.2f} ms of kernels per step, {sum(v[0] for v in tot.values()) // STEPS} launches per step")
GROUPS = [("batch norm (ours + ATen elementwise halves)", r"bn_|batch_norm"),
          ("kNN prepare / select / finalize", r"knn_|tc_prepare|tc::"),
          ("grouped FC fwd / dgrad / wgrad / pack", r"fc::|grouped_fc"),
          ("aggregate + key pooling", r"mr_aggregate|pool_keys"),
          ("label head / gather / loss", r"label_|neighbor_|multilabel"),
          ("library GEMMs / convolutions", r"nvjet|cutlass|xmma|splitK|gemm|cudnn|conv"),
          ("dtype / layout copies", r"copy"),
          ("optimizer / clip (multi-tensor)", r"multi_tensor|FusedOpt"),
          ("other ATen elementwise / reductions", r".")]
grp = collections.OrderedDict((g, [0, 0.0]) for g, _ in GROUPS)
for n, (c, t) in tot.items():
    for g, rx in GROUPS:
        if re.search(rx, n):
            grp[g][0] += c; grp[g][1] += t
            break
for g, (c, t) in grp.items():
    print(f"# {t / s * 100:5.1f}% {t / STEPS / 1e3:7.3f} ms x{c // STEPS:4d}  {g}")
for n, (c, t) in sorted(tot.items(), key=lambda x: -x[1][1])[:TOP]:
    print(f"{t / s * 100:5.1f}% {t / STEPS / 1e3:7.3f} ms x{c // STEPS:4d} avg {t / c:7.1f} us  {n}")
