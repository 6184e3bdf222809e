"""GPU: every kNN call of a whole GKGNet-576 forward (real, spatially smooth features under bf16 autocast -- unlike the
random activations of the layer tests) on the tcgen05 path returns the ids of the CUDA-core exact fp32 kernel, stays on
the fast path (a handful of rows on the brute-force fix-up at most) and within the certified error bound."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tc_delta(D):
    pa = (D + 2 + 15) // 16 * 16
    return 3.6e-7 * ((pa + (2 * D + 15) // 16 * 16) // 16) + 1.2e-6


def test_gkgnet576_every_knn_call_matches_exact_kernel():
    import gkgnet_b200 as G
    from gkgnet_b200 import _lib, ops

    calls = []
    orig = ops.knn_graph

    def wrapped(x, y=None, relative_pos=None, **kw):
        info = {"flags": 0}
        kw_tc = dict(kw)
        kw_tc.pop("debug", None)
        idx = orig(x, y, relative_pos, debug=info, **kw_tc)           # product path (AUTO) + counters
        kw_ex = dict(kw_tc)
        kw_ex.pop("separable", None)
        kw_ex["algo"] = _lib.KNN_EXACT_FP32
        ref = orig(x, y, relative_pos, **kw_ex)                       # exact kernel on the same features
        B, N, C = x.shape
        g = kw.get("groups", 1)
        calls.append(dict(N=N, M=N if y is None else y.shape[1], D=C // g, rows=B * g * N, dtype=x.dtype,
                          kd=kw.get("k", 9) * kw.get("dilation", 1), differ=int((idx != ref).any(-1).sum()),
                          **info["stats"]))
        return idx

    ops.knn_graph = wrapped
    G.set_norm_type("BN")
    try:
        torch.manual_seed(0)
        net = G.GKGNet(choice="s", n_classes=80, size=576, drop_path=0.0).cuda().eval()
        img = torch.randn(2, 3, 576, 576, device="cuda")
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            net(img)
    finally:
        ops.knn_graph = orig
        G.set_norm_type("SyncBN")
    assert len(calls) == 16, len(calls)                               # 12 Graphers + 4 label heads
    assert {c["D"] for c in calls} == {40, 80, 200, 320}
    for c in calls:
        assert c["differ"] == 0, c                                    # bit-equal neighbour ids
        assert c["fixups"] <= max(4, c["rows"] // 2000), c            # smooth features stay on the fast path
        assert c["max_err"] <= _tc_delta(c["D"]), c                   # observed |approx - exact| within the bound
