#!/usr/bin/env python
"""Native synchronised batch norm (layers.SyncBatchNorm) against torch.nn.SyncBatchNorm on N GPUs:
    torchrun --nproc-per-node 2 tools/syncbn_check.py
Every rank normalises its own batch with statistics over all ranks; outputs, input / parameter gradients and running
statistics must agree with the stock module (plain and fused with the GELU that follows)."""
import os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gkgnet_b200 import layers, parallel as P

world, rank, local = P.init_distributed()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
worst = 0.0
for C, H, dtype, act in ((160, 36, torch.bfloat16, False), (160, 36, torch.bfloat16, True), (80, 20, torch.float32, True),
                         (400, 18, torch.float32, False)):
    g = torch.Generator(device=dev).manual_seed(10 + rank)
    x = (torch.randn(4, C, H, H, device=dev, generator=g) * (1.0 + rank) + 0.5 * rank).to(dtype)
    x = x.contiguous(memory_format=torch.channels_last)
    dy = torch.randn(4, C, H, H, device=dev, generator=g).to(dtype).contiguous(memory_format=torch.channels_last)
    torch.manual_seed(0)
    ours, ref = layers.SyncBatchNorm(C).to(dev), torch.nn.SyncBatchNorm(C).to(dev)
    with torch.no_grad():
        ours.weight.uniform_(0.5, 1.5); ours.bias.normal_()
    ref.load_state_dict(ours.state_dict())
    gelu = torch.nn.GELU()
    res = []
    for m in (ours, ref):
        xi = x.clone().requires_grad_(True)
        if act:
            y = layers.run_modules([m, gelu], xi) if m is ours else gelu(m(xi))
        else:
            y = m(xi)
        y.backward(dy)
        res.append((y.detach().float(), xi.grad.float(), m.weight.grad.clone(), m.bias.grad.clone(),
                    m.running_mean.clone(), m.running_var.clone()))
    tol = 3e-2 if dtype == torch.bfloat16 else 1e-4
    names = ("y", "dx", "dweight", "dbias", "running_mean", "running_var")
    for n, a, b in zip(names, *res):
        scale = max(1.0, b.abs().max().item())
        err = (a - b).abs().max().item() / scale
        lim = tol * (10 if n in ("dweight", "dbias") and dtype == torch.bfloat16 else 1)
        assert err <= lim, (C, H, dtype, act, n, err)
        worst = max(worst, err)
    if rank == 0:
        print(f"C={C} H={H} {dtype} gelu={act}: ok")
dist.barrier()
if rank == 0:
    print(f"syncbn_check: world {world}, all cases agree with torch.nn.SyncBatchNorm (worst relative error {worst:.2e})")
dist.destroy_process_group()
