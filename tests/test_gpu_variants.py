"""GPU: the GraphConv2d variants beside max-relative (EdgeConv2d, GraphSAGE, GINConv2d, GraphAtten: torch_vertex.py:16-150)
on the native neighbour gather / sum kernels, against the reference's own outputs (tests/golden/gconv_*.npz), the
gather kernels against PyTorch indexing with autograd, and the stochastic branch of DenseDilated (torch_edge.py:139-146)."""
import pytest
import torch

from tests._util import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("conv", ["edge", "sage", "gin", "gat"])
def test_graphconv_variant_matches_reference(conv):
    import gkgnet_b200 as G
    g = load_golden("gconv_" + conv)
    G.set_norm_type("BN")
    try:
        m = G.GraphConv2d(16, 32, conv, "gelu", "batch", True)
    finally:
        G.set_norm_type("SyncBN")
    m.load_state_dict(g["sd"], strict=True)
    m = m.cuda().eval()
    with torch.no_grad():
        out = m(g["x"].cuda(), g["edge_index"].cuda(), g["y"].cuda())
        out_self = m(g["x"].cuda(), g["edge_index_self"].cuda(), None)
    assert tuple(out.shape) == tuple(g["out"].shape)
    assert torch.allclose(out.cpu(), g["out"], atol=1e-4, rtol=1e-4)
    assert torch.allclose(out_self.cpu(), g["out_self"], atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("dtype,groups", [(torch.float32, 1), (torch.float32, 2), (torch.bfloat16, 4)])
def test_neighbor_gather_and_sum_match_indexing(dtype, groups):
    from gkgnet_b200 import ops
    torch.manual_seed(1)
    B, N, M, C, k = 2, 50, 30, 24, 5
    D = C // groups
    y = torch.randn(B, M, C, device="cuda").to(dtype).requires_grad_(True)
    idx = torch.randint(0, M, (B * groups, N, k), device="cuda", dtype=torch.int32)
    got = ops.gather_neighbors(y, idx, groups=groups)
    s = ops.sum_neighbors(y, idx, groups=groups)
    # reference by indexing, group by group
    yr = y.detach().float().requires_grad_(True)
    parts = []
    for g in range(groups):
        ii = idx.view(B, groups, N, k)[:, g].long()                                   # (B, N, k)
        parts.append(torch.stack([yr[b, :, g * D:(g + 1) * D][ii[b]] for b in range(B)]))   # (B, N, k, D)
    want = torch.cat(parts, dim=-1)
    assert torch.equal(got.float(), want.to(dtype).float())
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert (s.float() - want.sum(2)).abs().max().item() <= tol * max(1.0, want.sum(2).abs().max().item())
    w1 = torch.randn_like(got)
    w2 = torch.randn_like(s)
    ((got * w1).sum() + (s * w2).sum()).backward()
    ((want * w1.float()).sum() + (want.sum(2) * w2.float()).sum()).backward()
    assert (y.grad.float() - yr.grad).abs().max().item() <= (1e-4 if dtype == torch.float32 else 5e-2) * max(1.0, yr.grad.abs().max().item())


def test_dynamic_graph_conv_with_variant_runs_end_to_end():
    """Grapher(conv='gin' / 'edge') without channel groups: forward + backward through kNN, gather and the convs."""
    import gkgnet_b200 as G
    G.set_norm_type("BN")
    try:
        for conv in ("gin", "edge", "sage", "gat"):
            m = G.Grapher(16, 4, 1, conv, "gelu", "batch", True, False, 0.2, 2, 64, 0.0, True, False, 1).cuda().train()
            x = torch.randn(2, 16, 8, 8, device="cuda", requires_grad=True)
            out = m(x)
            assert tuple(out.shape) == (2, 16, 8, 8)
            out.square().mean().backward()
            assert torch.isfinite(x.grad).all() and x.grad.abs().sum() > 0
    finally:
        G.set_norm_type("SyncBN")


def test_stochastic_dilation_picks_from_the_kd_nearest():
    """DenseDilated's stochastic branch (training, epsilon = 1): a random k of the k*d nearest (torch_edge.py:141-144)."""
    import gkgnet_b200 as G
    from gkgnet_b200 import ops
    torch.manual_seed(0)
    k, d = 4, 3
    g = G.DenseDilatedKnnGraph(k, d, stochastic=True, epsilon=1.0).cuda().train()
    x = torch.randn(2, 70, 16, device="cuda")
    picked = g.neighbors(x, None, None, groups=2)
    full = ops.knn_graph(x, None, None, groups=2, k=k * d, dilation=1)
    assert tuple(picked.shape) == (4, 70, k)
    inside = (picked.unsqueeze(-1) == full.unsqueeze(-2)).any(-1)
    assert inside.all()
    # the same k positions for every node (one randperm per call, like the reference), all distinct
    pos = (picked[0, 0].unsqueeze(-1) == full[0, 0].unsqueeze(0)).float().argmax(-1)
    assert len(set(pos.tolist())) == k
    assert torch.equal((picked.unsqueeze(-1) == full.unsqueeze(-2)).float().argmax(-1), pos.expand(4, 70, k))
    g.eval()
    assert torch.equal(g.neighbors(x, None, None, groups=2), full[:, :, ::d])
