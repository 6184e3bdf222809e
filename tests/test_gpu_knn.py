"""GPU parity: gkg_knn_graph (through the C ABI) against the CPU oracle.

Contract (BASELINE.json north_star): neighbour index sets bit-identical to the reference's
except at distance ties within 1e-6 relative.  `check_knn_against_distances` verifies, rank
by rank, that the distance of the returned neighbour equals the oracle's sorted distance
within rtol * max(1, |d|) and that ids are distinct -- i.e. any deviation is a tie."""
import pytest
import torch

from oracle import gkg_oracle as O
from tests._util import load_golden

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def _ops():
    import gkgnet_b200
    return gkgnet_b200.ops, gkgnet_b200._lib


def _ref_layout(t, G):
    """(B, N, C) token-major -> reference (B*G, D, N, 1)."""
    B, N, C = t.shape
    D = C // G
    return t.reshape(B, N, G, D).permute(0, 2, 3, 1).reshape(B * G, D, N, 1)


def _run_case(B, G, N, M, D, k, d, bias, algo, dtype=torch.float32, seed=0, quant=None, self_keys=False):
    ops, lib = _ops()
    if algo == lib.KNN_TCGEN05 and 3 * D + 2 > 640:
        algo = lib.KNN_AUTO          # D too large for the tensor-core tiling: AUTO falls back
    g = torch.Generator().manual_seed(seed)
    C = G * D
    x = torch.randn(B, N, C, generator=g)
    y = None if self_keys else torch.randn(B, M, C, generator=g)
    if quant:
        x = torch.round(x * quant) / quant
        y = None if y is None else torch.round(y * quant) / quant
    x = x.to(dtype)
    y = None if y is None else y.to(dtype)
    Mk = N if self_keys else M
    rel = -(0.5 + 0.5 * torch.rand(1, N, Mk, generator=g)) if bias else None
    idx = ops.knn_graph(x.cuda(), None if y is None else y.cuda(), None if rel is None else rel.cuda(),
                        groups=G, k=k, dilation=d, algo=algo)
    torch.cuda.synchronize()
    assert idx.dtype == torch.int32 and tuple(idx.shape) == (B * G, N, k)
    dist = O.knn_distance_matrix(_ref_layout(x.float(), G), None if y is None else _ref_layout(y.float(), G), rel)
    rep = O.check_knn_against_distances(idx.cpu(), dist, k, d, RTOL)
    assert rep["rows_bad"] == 0, rep
    return rep


ALGOS = ["exact", "tc"]


def _algo(name):
    _, lib = _ops()
    return {"exact": lib.KNN_EXACT_FP32, "auto": lib.KNN_AUTO, "tc": lib.KNN_TCGEN05}[name]


@pytest.mark.parametrize("algo", ALGOS)
def test_golden_knn_through_module(algo):
    import gkgnet_b200 as G
    for name in ("knn_xy", "knn_self"):
        g = load_golden(name)
        mod = G.DenseDilatedKnnGraph(g["k"], g["dilation"]).cuda()
        mod.algo = _algo(algo)
        y = g.get("y")
        ei = mod(g["x"].cuda(), None if y is None else y.cuda(), g["relative_pos"].cuda())
        assert ei.dtype == torch.int64 and tuple(ei.shape) == tuple(g["edge_index"].shape)
        assert torch.equal(ei.cpu()[1], g["edge_index"][1])
        dist = O.knn_distance_matrix(g["x"], y, g["relative_pos"])
        rep = O.check_knn_against_distances(ei.cpu()[0], dist, g["k"], g["dilation"], RTOL)
        assert rep["rows_bad"] == 0, rep
        # the fixture has no ties: must match the reference's own output exactly
        assert torch.equal(ei.cpu(), g["edge_index"])


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("B,G,N,M,D,k,d,bias", [
    (2, 2, 300, 70, 40, 9, 1, True),      # stage-1 like, ragged N/M
    (1, 2, 1296, 1296, 40, 9, 1, True),   # one full key set
    (2, 2, 80, 1000, 40, 9, 1, False),    # label head: few queries, many keys
    (2, 8, 150, 60, 10, 9, 1, True),      # G=8 sweep, D not a multiple of 4
    (1, 1, 130, 50, 17, 4, 2, False),     # odd D, single group
    (1, 4, 200, 64, 20, 18, 2, True),     # k=18 sweep (k*d = 36)
])
def test_knn_xy_random(algo, B, G, N, M, D, k, d, bias):
    _run_case(B, G, N, M, D, k, d, bias, _algo(algo))


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("B,G,N,D,k,d", [
    (2, 2, 324, 320, 9, 3),     # stage 4
    (1, 2, 1296, 200, 9, 3),    # stage 3, k*d = 27
    (2, 2, 100, 80, 9, 2),
    (1, 1, 324, 640, 9, 3),     # stage 4 with num_group = 1 (BASELINE configs[4] sweep)
    (1, 1, 784, 400, 9, 2),     # stage 3 at 448 px with num_group = 1
])
def test_knn_self_random(algo, B, G, N, D, k, d):
    rep = _run_case(B, G, N, None, D, k, d, True, _algo(algo), self_keys=True)
    assert rep["rows_differ"] <= rep["rows"] // 100


@pytest.mark.parametrize("algo", ALGOS)
def test_knn_bf16_inputs(algo):
    _run_case(2, 2, 400, 144, 40, 9, 1, True, _algo(algo), dtype=torch.bfloat16)
    _run_case(1, 2, 256, None, 200, 9, 2, True, _algo(algo), dtype=torch.bfloat16, self_keys=True)


@pytest.mark.parametrize("algo", ALGOS)
def test_knn_adversarial_ties(algo):
    # features on a coarse grid -> many exactly equal distances; every deviation from the
    # oracle's pick must still be a tie (rows_bad == 0 is asserted inside)
    rep = _run_case(1, 2, 256, 128, 8, 9, 1, False, _algo(algo), quant=2)
    assert rep["rows"] == 512


def test_knn_self_first_neighbor_is_self():
    ops, lib = _ops()
    x = torch.randn(2, 500, 80, device="cuda")
    for algo in (lib.KNN_EXACT_FP32, lib.KNN_AUTO):
        idx = ops.knn_graph(x, None, None, groups=2, k=9, dilation=1, algo=algo)
        want = torch.arange(500, device="cuda", dtype=torch.int32).expand(4, 500)
        assert torch.equal(idx[:, :, 0], want)


def test_knn_errors():
    ops, lib = _ops()
    x = torch.randn(1, 20, 8, device="cuda")
    with pytest.raises(RuntimeError):       # k*d > number of keys (torch.topk raises too)
        ops.knn_graph(x, None, None, groups=1, k=9, dilation=3)
    with pytest.raises(ValueError):
        ops.knn_graph(x, None, None, groups=3, k=2)
    with pytest.raises(RuntimeError):
        ops.knn_graph(x.cpu(), None, None, groups=1, k=2)
    empty = ops.knn_graph(torch.empty(0, 20, 8, device="cuda"), None, None, groups=2, k=3)
    assert tuple(empty.shape) == (0, 20, 3)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_knn_full_size_stage1_properties(dtype):
    """BASELINE config 2 shape (B=32, C=80, N=20736, M=1296, G=2, k=9), fp32 and the benchmarked bf16: too large
    for the CPU oracle in a test, so check size-independent properties: (a) the tcgen05 path (with and without the
    separable-bias hint) and the exact kernel agree up to ties, (b) on sampled rows the returned neighbours are the
    true top-k of an fp64 recomputation on the GPU (from the same rounded features), in ascending order, ties
    within 1e-6 relative only."""
    ops, lib = _ops()
    B, G, N, M, D, k = 32, 2, 20736, 1296, 40, 9
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, N, G * D, device="cuda", generator=g)
    y = torch.nn.functional.avg_pool2d(x.view(B, 144, 144, G * D).permute(0, 3, 1, 2), 4, 4)
    y = y.permute(0, 2, 3, 1).reshape(B, M, G * D).contiguous()
    x, y = x.to(dtype), y.to(dtype)
    from gkgnet_b200.pos_embed import relative_pos_table
    rel = relative_pos_table(G * D, N, 4)[0].cuda()
    sep = ops.fit_separable_bias(rel)
    assert sep is not None and sep[2:] == (144, 9)
    ia = ops.knn_graph(x, y, rel, groups=G, k=k, dilation=1, algo=lib.KNN_TCGEN05, separable=sep)
    idn = ops.knn_graph(x, y, rel, groups=G, k=k, dilation=1, algo=lib.KNN_TCGEN05)
    assert (ia != idn).any(-1).float().mean().item() < 1e-4
    ie = ops.knn_graph(x, y, rel, groups=G, k=k, dilation=1, algo=lib.KNN_EXACT_FP32)
    differ = (ia != ie).any(-1).float().mean().item()
    assert differ < 1e-3, differ
    rows = torch.randint(0, N, (256,), device="cuda", generator=g)
    for p in (0, 17, 63):
        b, gi = divmod(p, G)
        xs = torch.nn.functional.normalize(x[b, rows, gi * D:(gi + 1) * D].double(), dim=-1)
        ys = torch.nn.functional.normalize(y[b, :, gi * D:(gi + 1) * D].double(), dim=-1)
        dist = (xs * xs).sum(-1, keepdim=True) - 2 * xs @ ys.T + (ys * ys).sum(-1)[None] + rel[rows].double()
        for ids in (ia, ie):
            rep = O.check_knn_against_distances(ids[p, rows].cpu().unsqueeze(0), dist.cpu().unsqueeze(0), k, 1, 1e-6)
            assert rep["rows_bad"] == 0, rep
