#!/usr/bin/env python
"""Where does a GKGNet-576 step spend its GPU time?  torch.profiler kernel table for one train
(or inference) step; writes gpurun_out/model_profile_<mode>.txt."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gkgnet_b200 as G

mode = sys.argv[1] if len(sys.argv) > 1 else "train"
B = int(sys.argv[2]) if len(sys.argv) > 2 else (16 if mode == "train" else 64)
G.set_norm_type("BN")
dev = torch.device("cuda")
net = G.GKGNet(choice="s", n_classes=80, size=576, drop_path=0.1 if mode == "train" else 0.0).to(dev)
head = G.LabelQueryHead(80, 640).to(dev)
net.train(mode == "train"); head.train(mode == "train")
params = [p for p in list(net.parameters()) + list(head.parameters()) if p.requires_grad]
opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.05, fused=True)
img = torch.randn(B, 3, 576, 576, device=dev)
tgt = (torch.rand(B, 80, device=dev) < 0.04).float()

def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        if mode == "train":
            loss = sum(head.forward_train(net(img), tgt).values())
        else:
            with torch.no_grad():
                return torch.sigmoid(head.get_score(net(img)))
    loss.backward()
    torch.nn.utils.clip_grad_norm_(params, 5.0)
    opt.step(); opt.zero_grad(set_to_none=True)

for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
txt = prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", f"model_profile_{mode}.txt"), "w").write(txt)
print(txt[-6000:])
