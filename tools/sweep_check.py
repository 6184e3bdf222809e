#!/usr/bin/env python
"""BASELINE configs[4] sweep (k in {9,18} x channel groups in {1,4,8} x resolution {448,576}): does every
variant run through the CUDA path, and how fast is a forward?  (Also covered with parity at small sizes by tests.)"""
import itertools, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gkgnet_b200 as G

G.set_norm_type("BN")
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for size, k, g in itertools.product((448, 576), (9, 18), (1, 4, 8)):
    try:
        torch.manual_seed(0)
        net = G.GKGNet(choice="s", n_classes=80, size=size, k=k, k_label_gcn=k, num_group=g, drop_path=0.0).to(dev).eval()
        img = torch.randn(B, 3, size, size, device=dev)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            for _ in range(2):
                out = net(img)
            torch.cuda.synchronize()
            t0 = time.time()
            for _ in range(3):
                out = net(img)
            torch.cuda.synchronize()
        ms = (time.time() - t0) / 3 * 1e3
        ok = all(torch.isfinite(o.float()).all().item() for o in out[:2])
        print(f"size={size} k={k} groups={g}: ok={ok} {B / ms * 1e3:8.1f} img/s ({ms:.1f} ms / {B} images)", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"size={size} k={k} groups={g}: FAILED {type(e).__name__}: {str(e)[:160]}", flush=True)
    finally:
        net = None
        torch.cuda.empty_cache()
