"""GPU: bring-up and accuracy checks specific to the tcgen05 kNN kernel (debug hooks of
libgkg_b200.so: raw approximate distances out of TMEM, forced exact re-rank statistics)."""
import pytest
import torch

from oracle import gkg_oracle as O

pytestmark = pytest.mark.gpu

def tc_delta(D, dtype=torch.float32):
    """tc_delta() in csrc/knn_tc_kernel.cuh: certified |approx - exact| of the fp16x3 GEMM, distance units."""
    pa = (D + 2 + 15) // 16 * 16
    kp = pa + (2 * D + 15) // 16 * 16
    return 3.6e-7 * (kp // 16) + 1.2e-6


def _ref_layout(t, G):
    B, N, C = t.shape
    D = C // G
    return t.reshape(B, N, G, D).permute(0, 2, 3, 1).reshape(B * G, D, N, 1)



@pytest.mark.parametrize("B,G,N,M,D,bias,dtype", [
    (1, 2, 200, 300, 40, True, torch.float32),      # KP = 128, one k-block
    (2, 2, 128, 128, 40, False, torch.float32),     # exact tile multiples
    (1, 2, 260, 140, 80, True, torch.float32),      # KP = 256 -> 4 k-blocks of 64
    (1, 2, 200, 300, 40, True, torch.bfloat16),     # bf16 features
    (1, 2, 300, 300, 80, True, torch.bfloat16),     # 256-row items
    (1, 2, 150, 270, 200, True, torch.float32),     # split planes: hi resident, lo streamed, 7 + 7 k-blocks of 32
    (1, 2, 150, 270, 200, True, torch.bfloat16),
    (1, 2, 130, 330, 320, False, torch.bfloat16),   # stage-4 width: split planes, k-blocks of 32 (hi padded to 352)
    (1, 2, 140, 300, 96, True, torch.float32),      # narrowest split shape: k-blocks of 48, lo segment shorter than hi
    (1, 2, 140, 300, 120, False, torch.bfloat16),
    (1, 1, 130, 200, 400, False, torch.bfloat16),   # groups = 1 at stage 3: k-blocks of 16, 26 + 25 of them
    (1, 2, 100, 150, 100, True, torch.float32),     # D % 8 != 0: no split, operands streamed (K-concatenated rows)
    (1, 8, 100, 90, 10, False, torch.float32),      # KP = 32
    (1, 1, 90, 2000, 40, False, torch.bfloat16),    # label-head like: many key tiles
])
def test_tc_raw_distances(B, G, N, M, D, bias, dtype):
    from gkgnet_b200 import _lib, ops
    g = torch.Generator().manual_seed(7)
    C = G * D
    x = torch.randn(B, N, C, generator=g).to(dtype)
    y = torch.randn(B, M, C, generator=g).to(dtype)
    rel = -(0.5 + 0.5 * torch.rand(1, N, M, generator=g)) if bias else None
    dbg = torch.full((B * G, N, M), float("nan"), device="cuda")
    info = {"flags": 1, "dist": dbg}
    idx = ops.knn_graph(x.cuda(), y.cuda(), None if rel is None else rel.cuda(), groups=G, k=9,
                        dilation=1, algo=_lib.KNN_TCGEN05, debug=info)
    torch.cuda.synchronize()
    st = info["stats"]
    xr, yr = _ref_layout(x.float(), G), _ref_layout(y.float(), G)
    dist = O.knn_distance_matrix(xr, yr, rel)
    xn = O.l2_normalize(xr, 1).squeeze(-1)
    xsq = (xn * xn).sum(1)                      # (P, N)
    want = dist - xsq.unsqueeze(-1)             # the kernel ranks without the row constant
    err = (dbg.cpu() - want).abs().max().item()
    assert err < tc_delta(D, dtype) / 1.5, err
    assert st["max_err"] < tc_delta(D, dtype) / 1.5, st
    assert st["ambiguous"] == B * G * N          # forced re-rank touched every row
    rep = O.check_knn_against_distances(idx.cpu(), dist, 9, 1, 1e-6)
    assert rep["rows_bad"] == 0, rep


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_tc_matches_exact_bitwise_on_random_data(dtype):
    """After the exact re-rank of ambiguous rows the tcgen05 path must agree with the
    CUDA-core exact kernel except on true ties (random data: none)."""
    from gkgnet_b200 import _lib, ops
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(4, 2304, 80, device="cuda", generator=g).to(dtype)
    y = torch.randn(4, 576, 80, device="cuda", generator=g).to(dtype)
    rel = -(0.5 + 0.5 * torch.rand(2304, 576, device="cuda", generator=g))
    a = ops.knn_graph(x, y, rel, groups=2, k=9, dilation=1, algo=_lib.KNN_TCGEN05)
    b = ops.knn_graph(x, y, rel, groups=2, k=9, dilation=1, algo=_lib.KNN_EXACT_FP32)
    assert torch.equal(a, b)
    for ga in (18, 6, 3):       # keys per group of the threshold sweep: any choice must give the same neighbours
        a = ops.knn_graph(x, y, rel, groups=2, k=9, dilation=1, algo=_lib.KNN_TCGEN05, debug={"ga": ga})
        assert torch.equal(a, b), ga
    xs = x[:, :1296].contiguous()
    a = ops.knn_graph(xs, None, None, groups=2, k=9, dilation=3, algo=_lib.KNN_TCGEN05)
    b = ops.knn_graph(xs, None, None, groups=2, k=9, dilation=3, algo=_lib.KNN_EXACT_FP32)
    assert torch.equal(a, b)


@pytest.mark.parametrize("D,kd", [(200, 18), (200, 27), (320, 27), (40, 18), (80, 9), (96, 18), (200, 36)])
def test_tc_wide_groups_bf16_match_exact(D, kd):
    """Stage 3 / 4 shapes (self keys, wide groups, long lists) in bf16: ids equal the exact kernel's."""
    from gkgnet_b200 import _lib, ops
    g = torch.Generator(device="cuda").manual_seed(17)
    x = torch.randn(2, 1296, 2 * D, device="cuda", generator=g).to(torch.bfloat16)
    k, d = 9, kd // 9
    a = ops.knn_graph(x, None, None, groups=2, k=k, dilation=d, algo=_lib.KNN_TCGEN05)
    b = ops.knn_graph(x, None, None, groups=2, k=k, dilation=d, algo=_lib.KNN_EXACT_FP32)
    assert torch.equal(a, b)


@pytest.mark.parametrize("D,d", [(200, 2), (200, 3), (80, 1)])
def test_tc_several_items_per_cta(D, d):
    """More work items than SMs: the resident A tile is replaced between items, the operand ring and the three
    round-robin issuers carry their phases across items (B * G * ceil(N / 128) = 352 items on 148 CTAs)."""
    from gkgnet_b200 import _lib, ops
    from gkgnet_b200.pos_embed import relative_pos_table
    g = torch.Generator(device="cuda").manual_seed(23)
    x = torch.randn(16, 1296, 2 * D, device="cuda", generator=g).to(torch.bfloat16)
    rel = relative_pos_table(2 * D, 1296, 1).cuda()
    sep = ops.fit_separable_bias(rel)
    a = ops.knn_graph(x, None, rel, groups=2, k=9, dilation=d, algo=_lib.KNN_TCGEN05, separable=sep)
    b = ops.knn_graph(x, None, rel, groups=2, k=9, dilation=d, algo=_lib.KNN_EXACT_FP32)
    assert torch.equal(a, b)


@pytest.mark.parametrize("C,n,r,G,k,d", [
    (16, 81, 1, 2, 4, 2),        # 9x9 grid, self keys  -> Kw = 9
    (32, 324, 1, 2, 9, 3),       # 18x18 (stage-4 geometry) -> Kw = 18
    (80, 1296, 1, 2, 9, 2),      # 36x36 (stage-3 geometry) -> Kw = 36
    (80, 1296, 2, 2, 9, 1),      # 36x36 queries, 18x18 pooled keys -> M = 324, Kw = 9
    (160, 5184, 2, 2, 9, 1),     # stage-2 geometry: M = 1296, Kw = 18
])
def test_tc_separable_bias(C, n, r, G, k, d):
    """The analytic relative_pos table factorises; the register/shared-memory bias path must
    give the same raw distances as the dense table and the same neighbours as the oracle."""
    from gkgnet_b200 import _lib, ops
    from gkgnet_b200.pos_embed import relative_pos_table
    rel = relative_pos_table(C, n, r)
    side = int(n ** 0.5)
    g = torch.Generator().manual_seed(5)
    B = 1
    x = torch.randn(B, n, C, generator=g)
    y = None
    if r > 1:
        x4 = x.view(B, side, side, C).permute(0, 3, 1, 2)
        y = torch.nn.functional.avg_pool2d(x4, r, r).permute(0, 2, 3, 1).reshape(B, -1, C).contiguous()
    M = n // (r * r)
    sep = ops.fit_separable_bias(rel.cuda())
    assert sep is not None and sep[3] in (9, 18, 36), None if sep is None else sep[2:]
    dbg = torch.full((B * G, n, M), float("nan"), device="cuda")
    info = {"flags": 1, "dist": dbg}
    idx = ops.knn_graph(x.cuda(), None if y is None else y.cuda(), rel.cuda(), groups=G, k=k, dilation=d,
                        algo=_lib.KNN_TCGEN05, separable=sep, debug=info)
    torch.cuda.synchronize()
    st = info["stats"]
    xr = _ref_layout(x, G)
    yr = None if y is None else _ref_layout(y, G)
    dist = O.knn_distance_matrix(xr, yr, rel)
    xn = O.l2_normalize(xr, 1).squeeze(-1)
    want = dist - (xn * xn).sum(1).unsqueeze(-1)
    assert (dbg.cpu() - want).abs().max().item() < tc_delta(C // G) / 1.5
    assert st["max_err"] < tc_delta(C // G) / 1.5, st
    rep = O.check_knn_against_distances(idx.cpu(), dist, k, d, 1e-6)
    assert rep["rows_bad"] == 0, rep
    # and without the debug hooks, against the dense-table run
    a = ops.knn_graph(x.cuda(), None if y is None else y.cuda(), rel.cuda(), groups=G, k=k, dilation=d,
                      algo=_lib.KNN_TCGEN05, separable=sep)
    b = ops.knn_graph(x.cuda(), None if y is None else y.cuda(), rel.cuda(), groups=G, k=k, dilation=d,
                      algo=_lib.KNN_TCGEN05)
    assert torch.equal(a, b)


def test_fit_separable_rejects_arbitrary_tables():
    from gkgnet_b200 import ops
    rel = torch.rand(1, 81, 81, device="cuda")
    assert ops.fit_separable_bias(rel) is None


@pytest.mark.parametrize("N,M,D,bias", [(512, 1296, 40, True), (300, 648, 80, False)])
def test_tc_similar_neighbouring_keys_stay_on_fast_path(N, M, D, bias):
    """Image features are spatially smooth: consecutive keys (one logged triplet) are often all
    close to the query.  Such rows must be ranked by the kernel itself, not by the exact fix-up
    kernel (regression: the pair list overflowed and every row took the slow path)."""
    from gkgnet_b200 import _lib, ops
    g = torch.Generator().manual_seed(3)
    G, B = 2, 1
    C = G * D
    base = torch.randn(B, M // 3, C, generator=g)
    y = base.repeat_interleave(3, dim=1) + 2e-3 * torch.randn(B, M, C, generator=g)
    x = torch.randn(B, N, C, generator=g)
    rel = -(0.5 + 0.5 * torch.rand(1, N, M, generator=g)) if bias else None
    info = {"flags": 1}
    idx = ops.knn_graph(x.cuda(), y.cuda(), None if rel is None else rel.cuda(), groups=G, k=9,
                        dilation=1, algo=_lib.KNN_TCGEN05, debug=info)
    torch.cuda.synchronize()
    st = info["stats"]
    assert st["fixups"] <= 0.02 * B * G * N, st
    dist = O.knn_distance_matrix(_ref_layout(x, G), _ref_layout(y, G), rel)
    rep = O.check_knn_against_distances(idx.cpu(), dist, 9, 1, 1e-6)
    assert rep["rows_bad"] == 0, rep


def test_tc_smooth_key_field_compacts_instead_of_fixup():
    """Real feature maps are spatially smooth: the nearest keys of a query sit next to each other, i.e.
    in one or two 18-key groups of the threshold sweep, whose threshold is then far too loose.  The
    logging sweep must tighten it by compacting the row's log (regression: 20 % of the stage-1 rows of
    GKGNet-576 went to the brute-force fix-up kernel) and still return the exact neighbours."""
    from gkgnet_b200 import _lib, ops
    g = torch.Generator().manual_seed(9)
    B, G, D, side = 1, 2, 40, 36
    C, M, N = G * D, side * side, 1024
    # low-pass random field over the 36 x 36 key grid: neighbouring keys are nearly parallel
    f = torch.randn(B, C, side, side, generator=g)
    for _ in range(6):
        f = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(f, (1, 1, 1, 1), mode="replicate"), 3, 1)
    y = f.permute(0, 2, 3, 1).reshape(B, M, C).contiguous()
    pick = torch.randint(0, M, (N,), generator=g)
    x = (y[:, pick] + 0.05 * y.std() * torch.randn(B, N, C, generator=g)).contiguous()
    info = {"flags": 0}
    idx = ops.knn_graph(x.cuda(), y.cuda(), None, groups=G, k=9, dilation=1, algo=_lib.KNN_TCGEN05, debug=info)
    torch.cuda.synchronize()
    st = info["stats"]
    assert st["fixups"] <= 0.01 * B * G * N, st
    dist = O.knn_distance_matrix(_ref_layout(x, G), _ref_layout(y, G), None)
    rep = O.check_knn_against_distances(idx.cpu(), dist, 9, 1, 1e-6)
    assert rep["rows_bad"] == 0, rep


@pytest.mark.parametrize("B,G,N,M,D,k,d,keys", [
    (1, 2, 200, 300, 40, 9, 1, "random"),
    (1, 2, 150, 270, 200, 9, 3, "random"),      # k*d = 27
    (1, 1, 64, 2500, 40, 9, 2, "random"),       # more keys than candidate slots: the selected bin keeps few of them
    (1, 2, 96, 128, 8, 9, 2, "grid"),           # coarse grid: piles of exactly equal distances
    (1, 1, 48, 300, 16, 9, 2, "constant"),      # every distance ties: one bin, all keys are candidates
    (1, 1, 32, 2500, 16, 9, 1, "constant"),     # ... and more of them than candidate slots: count over all keys
    (1, 1, 24, 20736, 40, 9, 1, "random"),      # stage-1 label head: more keys than the shared-memory distance tile
                                                # (8192): distances in the global scratch row (M > 47104 limit)
    (1, 1, 8, 50176, 40, 9, 1, "random"),       # the 896-pixel label head the old kernel refused
    (1, 1, 40, 5184, 16, 9, 2, "few_unique"),   # fewer than k*d distinct near keys, thousands tied in the last
                                                # (clamped) histogram bin that holds the k*d-th neighbour
])
def test_tc_fixup_kernel_matches_exact_kernel(B, G, N, M, D, k, d, keys):
    """Debug hook 3 routes every row to the brute-force fix-up kernel (histogram select + rank by counting):
    its ids must equal the CUDA-core exact kernel's bit for bit, ties included (smaller key id first)."""
    from gkgnet_b200 import _lib, ops
    g = torch.Generator().manual_seed(13)
    C = G * D
    x = torch.randn(B, N, C, generator=g)
    y = torch.randn(B, M, C, generator=g)
    if keys == "grid":
        x, y = (x * 2).round() / 2, (y * 2).round() / 2
    elif keys == "constant":
        y = y[:, :1].expand(B, M, C).contiguous()
    elif keys == "few_unique":
        far = -x.mean(1, keepdim=True)                             # one far point, repeated
        y = torch.cat((y[:, :5], far.expand(B, M - 5, C)), 1).contiguous()
    x, y = x.cuda(), y.cuda()
    info = {"flags": 3}
    got = ops.knn_graph(x, y, None, groups=G, k=k, dilation=d, algo=_lib.KNN_TCGEN05, debug=info)
    torch.cuda.synchronize()
    st = info["stats"]
    assert st["fixups"] == B * G * N, st
    want = ops.knn_graph(x, y, None, groups=G, k=k, dilation=d, algo=_lib.KNN_EXACT_FP32)
    assert torch.equal(got, want)
    if keys == "constant":
        assert torch.equal(got[0, 0].cpu(), torch.arange(0, k * d, d, dtype=torch.int32))
