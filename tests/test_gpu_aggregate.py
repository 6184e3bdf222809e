"""GPU parity: max-relative aggregation forward/backward (C ABI) against the CPU oracle.
Tolerances from BASELINE.json: 1e-4 (fp32), 2e-2 (bf16); forward is in fact bit-exact."""
import pytest
import torch

from oracle import gkg_oracle as O
from tests._util import load_golden

pytestmark = pytest.mark.gpu


def _ref_layout(t, G):
    B, N, C = t.shape
    D = C // G
    return t.reshape(B, N, G, D).permute(0, 2, 3, 1).reshape(B * G, D, N, 1)


def _oracle_agg(x, idx, y, G):
    """token-major in/out wrapper around oracle.mr_aggregate."""
    B, N, C = x.shape
    P = B * G
    center = torch.arange(N).view(1, N, 1).expand(P, N, idx.shape[-1])
    ei = torch.stack((idx.long(), center), 0)
    out = O.mr_aggregate(_ref_layout(x, G), ei, None if y is None else _ref_layout(y, G), in_channels=C)
    return out.squeeze(-1).transpose(1, 2)          # (B, N, 2C)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,G,N,M,D,k", [
    (2, 2, 300, 70, 40, 9),
    (1, 2, 1296, 324, 80, 9),
    (2, 8, 150, 60, 10, 9),      # D=10: 4-byte vector path for bf16
    (1, 1, 33, 17, 7, 3),        # odd D: scalar path
    (2, 2, 80, 500, 200, 18),
    (3, 2, 1100, 300, 16, 9),    # bf16: keys staged in shared memory, CTA ranges crossing images
    (2, 2, 1500, 400, 40, 18),   # bf16: shared-memory path, k = 18
    (1, 2, 1296, 1296, 200, 9),  # stage-3 geometry: 10 channel slices of 40
    (1, 2, 1030, 520, 200, 9),   # 40-channel slices, six nodes per warp pass, ragged last pass (1030 % 6 = 4)
    (2, 2, 1027, 300, 80, 18),   # 80-channel slices of one group, k = 18, ragged last pass, ranges crossing images
])
def test_aggregate_forward(dtype, B, G, N, M, D, k):
    from gkgnet_b200 import ops
    g = torch.Generator().manual_seed(1)
    C = G * D
    x = torch.randn(B, N, C, generator=g).to(dtype)
    y = torch.randn(B, M, C, generator=g).to(dtype)
    idx = torch.randint(0, M, (B * G, N, k), generator=g, dtype=torch.int32)
    out = ops.mr_aggregate(x.cuda(), idx.cuda(), y.cuda(), groups=G)
    want = _oracle_agg(x.float(), idx, y.float(), G).to(dtype)
    assert out.dtype == dtype and tuple(out.shape) == (B, N, 2 * C)
    assert torch.equal(out.cpu(), want)
    # self keys
    idx_s = torch.randint(0, N, (B * G, N, k), generator=g, dtype=torch.int32)
    out_s = ops.mr_aggregate(x.cuda(), idx_s.cuda(), None, groups=G)
    assert torch.equal(out_s.cpu(), _oracle_agg(x.float(), idx_s, None, G).to(dtype))


def test_aggregate_strided_channels_last_view():
    """x given as a channels_last NCHW tensor viewed token-major, y as a strided slice."""
    from gkgnet_b200 import ops
    g = torch.Generator().manual_seed(2)
    B, C, H, W, G, k = 2, 16, 6, 5, 2, 4
    x4 = torch.randn(B, C, H, W, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    xt = x4.permute(0, 2, 3, 1).reshape(B, H * W, C)
    assert xt.data_ptr() == x4.data_ptr()
    ybig = torch.randn(B, 20, 2 * C, generator=g).cuda()
    yt = ybig[:, :, :C]                      # row stride 2C, still channel-contiguous
    idx = torch.randint(0, 20, (B * G, H * W, k), generator=g, dtype=torch.int32)
    out = ops.mr_aggregate(xt, idx.cuda(), yt, groups=G)
    want = _oracle_agg(xt.cpu(), idx, yt.cpu().contiguous(), G)
    assert torch.equal(out.cpu(), want)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("self_keys", [False, True])
def test_aggregate_backward(dtype, tol, self_keys):
    from gkgnet_b200 import ops
    g = torch.Generator().manual_seed(3)
    B, G, N, M, D, k = 2, 2, 200, 48, 40, 9
    C = G * D
    Mk = N if self_keys else M
    x = torch.randn(B, N, C, generator=g).to(dtype)
    y = None if self_keys else torch.randn(B, M, C, generator=g).to(dtype)
    idx = torch.randint(0, Mk, (B * G, N, k), generator=g, dtype=torch.int32)
    w = torch.randn(B, N, 2 * C, generator=g).to(dtype)

    xc = x.cuda().requires_grad_(True)
    yc = None if y is None else y.cuda().requires_grad_(True)
    out = ops.mr_aggregate(xc, idx.cuda(), yc, groups=G)
    out.backward(w.cuda())

    xo = x.float().requires_grad_(True)
    yo = None if y is None else y.float().requires_grad_(True)
    (_oracle_agg(xo, idx, yo, G) * w.float()).sum().backward()
    scale = max(1.0, float(xo.grad.abs().max()))
    assert (xc.grad.float().cpu() - xo.grad).abs().max() <= tol * scale
    if not self_keys:
        scale = max(1.0, float(yo.grad.abs().max()))
        assert (yc.grad.float().cpu() - yo.grad).abs().max() <= tol * scale


@pytest.mark.parametrize("quantised", [False, True])
@pytest.mark.parametrize("B,G,N,M,D,k,self_keys", [
    (2, 2, 1100, 300, 40, 9, False),     # warp-autonomous kernel: 80-channel slices holding both groups
    (2, 2, 1100, 300, 40, 18, False),    # k = 18: two id pairs per lane
    (1, 2, 1200, 300, 80, 9, False),     # 80-channel slices, one group per slice
    (1, 2, 1296, 1296, 200, 9, True),    # stage-3 geometry: 40-channel slices, self keys
    (3, 2, 1100, 300, 16, 9, False),     # CTA-tiled shared-memory kernel
])
def test_aggregate_backward_large(B, G, N, M, D, k, self_keys, quantised):
    """Backward through the bf16 shared-memory forward kernels (N >= 1024): the arg-max plane they write
    routes grad_y.  With features quantised to steps of 1/4 most maxima are tied: the smallest list
    position must win, like torch.max on the CPU (first maximal value)."""
    from gkgnet_b200 import ops
    g = torch.Generator().manual_seed(5)
    C = G * D
    dtype, tol = torch.bfloat16, 2e-2
    q = (lambda t: (t * 4).round() / 4) if quantised else (lambda t: t)
    x = q(torch.randn(B, N, C, generator=g)).to(dtype)
    y = None if self_keys else q(torch.randn(B, M, C, generator=g)).to(dtype)
    idx = torch.randint(0, N if self_keys else M, (B * G, N, k), generator=g, dtype=torch.int32)
    w = torch.randn(B, N, 2 * C, generator=g).to(dtype)

    xc = x.cuda().requires_grad_(True)
    yc = None if y is None else y.cuda().requires_grad_(True)
    out = ops.mr_aggregate(xc, idx.cuda(), yc, groups=G)
    out.backward(w.cuda())

    xo = x.float().requires_grad_(True)
    yo = None if y is None else y.float().requires_grad_(True)
    oo = _oracle_agg(xo, idx, yo, G)
    assert torch.equal(out.detach().cpu(), oo.detach().to(dtype))
    (oo * w.float()).sum().backward()
    scale = max(1.0, float(xo.grad.abs().max()))
    assert (xc.grad.float().cpu() - xo.grad).abs().max() <= tol * scale
    if not self_keys:
        scale = max(1.0, float(yo.grad.abs().max()))
        assert (yc.grad.float().cpu() - yo.grad).abs().max() <= tol * scale


def test_aggregate_many_keys_fall_back():
    """More than 65535 keys: ids no longer fit the 16-bit hand-over of the warp-autonomous kernel."""
    from gkgnet_b200 import ops
    g = torch.Generator().manual_seed(6)
    B, G, N, D, k = 1, 2, 70000, 40, 9
    C = G * D
    x = torch.randn(B, N, C, generator=g).bfloat16()
    idx = torch.randint(0, N, (B * G, N, k), generator=g, dtype=torch.int32)
    idx[:, :, 0] = N - 1 - torch.arange(N, dtype=torch.int32)        # ids above 65535 in use
    out = ops.mr_aggregate(x.cuda(), idx.cuda(), None, groups=G)
    assert torch.equal(out.cpu(), _oracle_agg(x.float(), idx, None, G).bfloat16())


def test_golden_mrconv_module():
    """Reference-layout MRConv2d.forward(x, edge_index, y) with the reference's weights."""
    import gkgnet_b200 as G
    g = load_golden("mrconv")
    m = G.MRConv2d(16, 32, "gelu", "batch", True)
    m.load_state_dict(g["sd"])
    m = m.cuda().eval()
    out = m(g["x"].cuda(), g["edge_index"].cuda(), g["y"].cuda())
    assert tuple(out.shape) == tuple(g["out"].shape)
    assert torch.allclose(out.cpu(), g["out"], atol=1e-4, rtol=1e-4)
    ei_self = O.dense_dilated_knn_graph(g["x"], None, 3, 1, None)
    out_s = m(g["x"].cuda(), ei_self.cuda(), None)
    assert torch.allclose(out_s.cpu(), g["out_self"], atol=1e-4, rtol=1e-4)


def test_aggregate_full_size_stage1_linearity():
    """BASELINE config 2 shape, bf16.  Size-independent checks: (a) even channels reproduce
    x bit-exactly, (b) odd channels + x equal the max over gathered rows recomputed with
    torch on the GPU, (c) backward conserves mass: sum(grad_x) + sum(grad_y) == sum over even
    grad_out channels (the -1 and +1 routes of the max-relative term cancel)."""
    from gkgnet_b200 import ops
    B, G, N, M, D, k = 32, 2, 20736, 1296, 40, 9
    C = G * D
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, N, C, device="cuda", generator=g).bfloat16().requires_grad_(True)
    y = torch.randn(B, M, C, device="cuda", generator=g).bfloat16().requires_grad_(True)
    idx = torch.randint(0, M, (B * G, N, k), device="cuda", generator=g, dtype=torch.int32)
    out = ops.mr_aggregate(x, idx, y, groups=G)
    assert torch.equal(out[..., 0::2], x.detach())
    b = 5
    for gi in range(G):
        rows = y.detach()[b, :, gi * D:(gi + 1) * D][idx[b * G + gi].long()]      # (N, k, D)
        want = (rows.float().max(1).values - x.detach()[b, :, gi * D:(gi + 1) * D].float()).bfloat16()
        assert torch.equal(out[b, :, 1::2][:, gi * D:(gi + 1) * D], want)
    go = torch.randn(B, N, 2 * C, device="cuda", generator=g).bfloat16()
    out.backward(go)
    lhs = x.grad.double().sum() + y.grad.double().sum()
    rhs = go[..., 0::2].double().sum()
    assert abs(float(lhs - rhs)) < 2e-2 * float(go.double().abs().sum()) ** 0.5 + 64


def test_aggregate_backward_deterministic_option():
    """The 64-bit fixed-point scatter (ops.set_deterministic_aggregate / torch.use_deterministic_algorithms): two runs
    agree bit for bit, and with the fp32-atomic path up to fp32 summation order."""
    from gkgnet_b200 import ops
    g = torch.Generator().manual_seed(9)
    B, G, N, M, D, k = 2, 2, 3000, 200, 40, 9
    C = G * D
    x = torch.randn(B, N, C, generator=g).bfloat16().cuda().requires_grad_(True)
    y = torch.randn(B, M, C, generator=g).bfloat16().cuda().requires_grad_(True)
    idx = torch.randint(0, M, (B * G, N, k), generator=g, dtype=torch.int32).cuda()
    w = torch.randn(B, N, 2 * C, generator=g).bfloat16().cuda()

    def grads():
        x.grad = y.grad = None
        ops.mr_aggregate(x, idx, y, groups=G).backward(w)
        return x.grad.clone(), y.grad.clone()

    plain = grads()
    ops.set_deterministic_aggregate(True)
    try:
        a, b = grads(), grads()
    finally:
        ops.set_deterministic_aggregate(None)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert torch.equal(a[0], plain[0])
    scale = max(1.0, plain[1].float().abs().max().item())
    assert (a[1].float() - plain[1].float()).abs().max().item() <= 8e-3 * scale      # one bf16 ulp of the largest sum


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,G,N,M,D,k", [
    (2, 2, 300, 70, 40, 9),       # generic / CTA-tiled kernels
    (2, 2, 1500, 400, 40, 9),     # bf16: warp-autonomous shared-memory kernel
    (1, 1, 33, 17, 7, 3),         # scalar path
])
def test_aggregate_propagates_nan_like_torch_max(dtype, B, G, N, M, D, k):
    """torch.max(x_j - x_i, -1) propagates NaN (torch_vertex.py:53): a NaN key must poison exactly the (node, channel)
    outputs whose neighbour list contains it, in every kernel variant (the kernels used to drop NaNs)."""
    from gkgnet_b200 import ops
    g = torch.Generator().manual_seed(4)
    C = G * D
    x = torch.randn(B, N, C, generator=g).to(dtype)
    y = torch.randn(B, M, C, generator=g).to(dtype)
    y[0, 3, 1] = float("nan")
    y[B - 1, M - 1, C - 1] = float("nan")
    idx = torch.randint(0, M, (B * G, N, k), generator=g, dtype=torch.int32)
    out = ops.mr_aggregate(x.cuda(), idx.cuda(), y.cuda(), groups=G).cpu()
    want = _oracle_agg(x.float(), idx, y.float(), G).to(dtype)
    assert torch.isnan(want).any()
    assert torch.equal(torch.isnan(out), torch.isnan(want))
    ok = ~torch.isnan(want)
    assert torch.equal(out[ok], want[ok])
