import sys, time, torch
sys.path.insert(0, '/root/repo')
from gkgnet_b200 import ops, _lib
torch.manual_seed(0)
for (B, G, N, M, D) in [(64, 2, 80, 20736, 40), (64, 2, 80, 5184, 80), (64, 2, 80, 1296, 200), (64, 2, 80, 324, 320), (64, 2, 1296, 1296, 200)]:
    x = torch.randn(B, N, G * D, device='cuda').to(torch.bfloat16)
    y = torch.randn(B, M, G * D, device='cuda').to(torch.bfloat16)
    for algo, name in ((_lib.KNN_TCGEN05, 'tc'), (_lib.KNN_EXACT_FP32, 'exact')):
        for _ in range(2):
            ops.knn_graph(x, y, None, groups=G, k=9, dilation=1, algo=algo)
        torch.cuda.synchronize(); t0 = time.time()
        for _ in range(5):
            ops.knn_graph(x, y, None, groups=G, k=9, dilation=1, algo=algo)
        torch.cuda.synchronize()
        print(f"N={N} M={M} D={D} {name}: {(time.time()-t0)/5*1e3:.3f} ms", flush=True)
