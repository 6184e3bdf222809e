#!/usr/bin/env python
"""Debug: per-key-tile clock64 stamps of CTA 0 of the tcgen05 kNN kernel on the bench shape.
Slots per tile: 0 MMA before t_empty wait, 1 after, 2 after commit(t_full), 3 epilogue (warp 2) before
t_full wait, 4 after, 5 at release, 6 producer issued the tile's last block.  Writes gpurun_out/knn_trace.json."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from gkgnet_b200 import _lib, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
lib = _lib.load()
if os.environ.get("GKG_LABEL"):          # label-head shape: 80 queries against all stage-1 patches, no bias
    torch.manual_seed(0)
    x = torch.randn(B, 80, 80).to(torch.bfloat16).cuda()
    y = torch.randn(B, 20736, 80).to(torch.bfloat16).cuda()
    rel, sep = None, None
else:
    x, y, rel = bench.make_inputs(B, "cpu", torch.bfloat16, 0)
    x, y, rel = x.cuda(), y.cuda(), rel.cuda()
    sep = ops.fit_separable_bias(rel)
for _ in range(2):
    ops.knn_graph(x, y, rel, groups=2, k=9, dilation=1, algo=_lib.KNN_TCGEN05, separable=sep)
torch.cuda.synchronize()
KT = int(sys.argv[2]) if len(sys.argv) > 2 else 18
tiles = (1 if os.environ.get('GKG_LABEL') else 36) * KT
buf = torch.zeros(tiles * 8, dtype=torch.int64, device="cuda")
lib.gkg_debug_knn_tc_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.gkg_debug_knn_tc_trace(buf.data_ptr(), tiles)
if os.environ.get("GKG_SKIP_PROCESS"):
    lib.gkg_debug_knn_tc.argtypes = [ctypes.c_int, ctypes.c_void_p]
    lib.gkg_debug_knn_tc(-2, None)
ops.knn_graph(x, y, rel, groups=2, k=9, dilation=1, algo=_lib.KNN_TCGEN05, separable=sep)
torch.cuda.synchronize()
lib.gkg_debug_knn_tc_trace(None, 0)
t = buf.view(tiles, 8).cpu()
t0 = int(t[0, 0])
rows = (t - t0).tolist()
for i, r in enumerate(rows):
    r[7] = int(t[i, 7]); r[6] = int(t[i, 6])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "knn_trace.json"), "w"))
import statistics as st
def col(i): return [r[i] for r in rows]
mma_wait = [r[1] - r[0] for r in rows]; mma_work = [r[2] - r[1] for r in rows]
epi_wait = [r[4] - r[3] for r in rows]
print("total cycles", rows[-1][5], "tiles", tiles)
print("mma  t_empty wait mean", st.mean(mma_wait), "work mean", st.mean(mma_work))
print("epi  t_full wait mean", st.mean(epi_wait))
for kt in range(KT):
    sel = [i for i in range(tiles) if i % KT == kt and i >= KT]
    print(kt, "epi wait", int(st.mean(epi_wait[i] for i in sel)), "tile time", int(st.mean(rows[i][5] - rows[i - 1][5] for i in sel)),
          "full->epi lag", int(st.mean(rows[i][4] - rows[i][2] for i in sel)), "mma wait", int(st.mean(mma_wait[i] for i in sel)),
          "mma work", int(st.mean(mma_work[i] for i in sel)), "of which b_full wait", int(st.mean(rows[i][7] for i in sel)), "issue", int(st.mean(rows[i][6] for i in sel)), "release->mma lag", int(st.mean(rows[i][1] - rows[i - 3][5] for i in sel)))
for r in rows[2 * KT:4 * KT][:8]:
    print(r)
