import sys, torch
sys.path.insert(0, '/root/repo')
from gkgnet_b200 import ops, _lib
torch.manual_seed(0)
B, G, N, M, D = 64, 2, 80, 20736, 40
x = torch.randn(B, N, G * D, device='cuda').to(torch.bfloat16)
y = torch.randn(B, M, G * D, device='cuda').to(torch.bfloat16)
for _ in range(3):
    ops.knn_graph(x, y, None, groups=G, k=9, dilation=1, algo=_lib.KNN_TCGEN05)
torch.cuda.synchronize()
