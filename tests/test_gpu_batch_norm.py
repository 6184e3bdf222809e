"""Training-mode batch norm: gkg_bn_stats / gkg_bn_backward_reduce behind ops.batch_norm_train
and layers.BatchNorm2d, against ATen's F.batch_norm (what `norm_layer('batch')`, torch_nn.py:32-42, resolves to)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("B,C,H,W,dtype", [
    (4, 80, 36, 36, torch.bfloat16),
    (2, 160, 20, 21, torch.bfloat16),        # ragged row count
    (16, 2560, 18, 18, torch.bfloat16),      # 4C at stage 4: two column slabs
    (3, 40, 7, 9, torch.bfloat16),           # stem width, few rows
    (2, 400, 36, 36, torch.float32),
    (5, 12, 11, 13, torch.float32),          # C % 4 == 0 only
    (16, 160, 144, 144, torch.bfloat16),     # the stage-1 BasicConv norm of the training benchmark
])
def test_batch_norm_train_matches_aten(B, C, H, W, dtype):
    from gkgnet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    x = _cl((torch.randn(B, C, H, W, device="cuda", generator=g) * 1.7 + 0.3).to(dtype))
    dy = _cl(torch.randn(B, C, H, W, device="cuda", generator=g).to(dtype))
    w = torch.rand(C, device="cuda", generator=g) + 0.5
    b = torch.randn(C, device="cuda", generator=g)
    res = []
    for native in (True, False):
        xi = x.clone().requires_grad_(True)
        wi, bi = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
        if native:
            y = ops.batch_norm_train(xi, wi, bi, rm, rv, 0.1, 1e-5)
        else:
            y = F.batch_norm(xi, rm, rv, wi, bi, True, 0.1, 1e-5)
        y.backward(dy)
        res.append((y.detach().float(), xi.grad.float(), wi.grad, bi.grad, rm, rv))
    (y0, dx0, dw0, db0, rm0, rv0), (y1, dx1, dw1, db1, rm1, rv1) = res
    tol = 2e-2 if dtype == torch.bfloat16 else 2e-5
    assert y0.shape == y1.shape and y0.dtype == y1.dtype
    assert torch.allclose(y0, y1, atol=tol, rtol=tol)
    assert torch.allclose(dx0, dx1, atol=tol, rtol=tol)
    rows = B * H * W
    # parameter gradients are sums over the rows: compare relative to their scale
    assert (dw0 - dw1).abs().max() <= 1e-3 * max(1.0, dw1.abs().max().item()) * (10 if dtype == torch.bfloat16 else 1)
    assert (db0 - db1).abs().max() <= 1e-3 * max(1.0, db1.abs().max().item())
    assert torch.allclose(rm0, rm1, atol=1e-5, rtol=1e-5)
    assert torch.allclose(rv0, rv1, atol=1e-5, rtol=1e-4)
    assert rows > 1


def test_batch_norm_stats_large_mean_fp64():
    """|mean| >> std: the shifted block sums + Chan merge must not lose the variance (naive sum / sum-of-squares in
    fp32 would: 1e4^2 * 2^-24 ~ 6 > var)."""
    from gkgnet_b200 import _lib, ops
    g = torch.Generator(device="cuda").manual_seed(9)
    C = 64
    x = _cl(torch.randn(8, C, 60, 60, device="cuda", generator=g) * 0.5 + 1.0e4)
    rows = x.shape[0] * x.shape[2] * x.shape[3]
    lib = _lib.load()
    mean = torch.empty(C, device="cuda")
    invstd = torch.empty(C, device="cuda")
    ws = torch.empty(lib.gkg_bn_workspace_bytes(rows, C), dtype=torch.uint8, device="cuda")
    x2 = x.permute(0, 2, 3, 1).reshape(rows, C)
    rc = lib.gkg_bn_stats(x2.data_ptr(), rows, C, 0, 1e-5, 0.1, mean.data_ptr(), invstd.data_ptr(), None, None,
                          ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    xd = x2.double()
    want_mean = xd.mean(0)
    want_invstd = (xd.var(0, unbiased=False) + 1e-5).rsqrt()
    assert torch.allclose(mean.double(), want_mean, rtol=5e-7, atol=0)      # a few fp32 ulps at 1e4
    assert torch.allclose(invstd.double(), want_invstd, rtol=2e-3)


def test_batch_norm_is_deterministic():
    from gkgnet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    x = _cl(torch.randn(8, 320, 72, 72, device="cuda", generator=g).to(torch.bfloat16))
    dy = _cl(torch.randn(8, 320, 72, 72, device="cuda", generator=g).to(torch.bfloat16))
    outs = []
    for _ in range(2):
        xi = x.clone().requires_grad_(True)
        w = torch.ones(320, device="cuda", requires_grad=True)
        b = torch.zeros(320, device="cuda", requires_grad=True)
        y = ops.batch_norm_train(xi, w, b, None, None, 0.1, 1e-5)
        y.backward(dy)
        outs.append((y.detach(), xi.grad, w.grad, b.grad))
    for a, c in zip(*outs):
        assert torch.equal(a, c)


def test_batch_norm_module_keeps_state_dict_and_falls_back():
    from gkgnet_b200.layers import BatchNorm2d
    torch.manual_seed(0)
    ours, ref = BatchNorm2d(80).cuda(), torch.nn.BatchNorm2d(80).cuda()
    ref.load_state_dict(ours.state_dict())
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    x = torch.randn(4, 80, 24, 24, device="cuda")
    for inp in (_cl(x), x.contiguous()):            # channels-last -> native kernels; NCHW -> stock module
        ya, yb = ours(inp), ref(inp)
        assert torch.allclose(ya, yb, atol=2e-5, rtol=2e-5)
    assert int(ours.num_batches_tracked) == int(ref.num_batches_tracked) == 2
    assert torch.allclose(ours.running_mean, ref.running_mean, atol=1e-6)
    assert torch.allclose(ours.running_var, ref.running_var, rtol=1e-5)
    ours.eval(), ref.eval()
    assert torch.allclose(ours(_cl(x)), ref(_cl(x)), atol=1e-6)
    with pytest.raises(ValueError):
        from gkgnet_b200 import ops
        ops.batch_norm_train(x.contiguous(), ours.weight, ours.bias, None, None, 0.1, 1e-5)   # not channels-last


@pytest.mark.parametrize("rows,C,dtype", [(20736, 160, torch.bfloat16), (1001, 2560, torch.bfloat16), (777, 12, torch.float32)])
def test_column_sum(rows, C, dtype):
    from gkgnet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(rows, C, device="cuda", generator=g).to(dtype)
    got = ops.column_sum(x)
    want = x.double().sum(0)
    assert got.dtype == torch.float32
    assert torch.allclose(got.double(), want, rtol=1e-5, atol=1e-3)
    assert torch.equal(got, ops.column_sum(x))          # deterministic


def test_conv1x1_matches_conv2d_with_autograd():
    """layers.FoldedSequential routes 1x1 convolutions on channels-last activations through ops.conv1x1."""
    from gkgnet_b200 import ops
    torch.manual_seed(1)
    conv = torch.nn.Conv2d(80, 160, 1).cuda()
    x = _cl(torch.randn(4, 80, 20, 20, device="cuda"))
    dy = _cl(torch.randn(4, 160, 20, 20, device="cuda"))
    xa = x.clone().requires_grad_(True)
    ya = ops.conv1x1(xa, conv.weight, conv.bias)
    ya.backward(dy)
    ga = (xa.grad.clone(), conv.weight.grad.clone(), conv.bias.grad.clone())
    conv.zero_grad()
    xb = x.clone().requires_grad_(True)
    yb = conv(xb)
    yb.backward(dy)
    assert torch.allclose(ya, yb, atol=1e-4, rtol=1e-4)
    for a, b in zip(ga, (xb.grad, conv.weight.grad, conv.bias.grad)):
        assert a.shape == b.shape and a.dtype == b.dtype
        assert torch.allclose(a, b, atol=2e-3, rtol=1e-3)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        yc = ops.conv1x1(x.to(torch.bfloat16), conv.weight, conv.bias)
    assert yc.dtype == torch.bfloat16 and torch.allclose(yc.float(), yb, atol=5e-2, rtol=5e-2)


@pytest.mark.parametrize("B,C,H,W,dtype", [
    (4, 160, 36, 36, torch.bfloat16),
    (2, 320, 20, 21, torch.bfloat16),
    (3, 40, 17, 9, torch.bfloat16),
    (2, 400, 18, 18, torch.float32),
])
def test_batch_norm_gelu_fused_matches_aten(B, C, H, W, dtype):
    """norm -> nn.GELU() as one forward pass and two backward passes (gkg_bn_act_forward / _backward)."""
    from gkgnet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(6)
    x = _cl((torch.randn(B, C, H, W, device="cuda", generator=g) * 1.3 - 0.2).to(dtype))
    dy = _cl(torch.randn(B, C, H, W, device="cuda", generator=g).to(dtype))
    w = torch.rand(C, device="cuda", generator=g) + 0.5
    b = torch.randn(C, device="cuda", generator=g) * 0.5
    res = []
    for native in (True, False):
        xi = x.clone().requires_grad_(True)
        wi, bi = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
        if native:
            y = ops.batch_norm_train(xi, wi, bi, rm, rv, 0.1, 1e-5, act="gelu")
        else:
            y = F.gelu(F.batch_norm(xi, rm, rv, wi, bi, True, 0.1, 1e-5))
        y.backward(dy)
        res.append((y.detach().float(), xi.grad.float(), wi.grad, bi.grad, rm, rv))
    (y0, dx0, dw0, db0, rm0, rv0), (y1, dx1, dw1, db1, rm1, rv1) = res
    tol = 2e-2 if dtype == torch.bfloat16 else 3e-5
    assert torch.allclose(y0, y1, atol=tol, rtol=tol)
    assert torch.allclose(dx0, dx1, atol=tol, rtol=tol)
    scale = 10 if dtype == torch.bfloat16 else 1       # the unfused bf16 path rounds the norm output and dgelu to bf16
    assert (dw0 - dw1).abs().max() <= 2e-3 * scale * max(1.0, dw1.abs().max().item())
    assert (db0 - db1).abs().max() <= 2e-3 * scale * max(1.0, db1.abs().max().item())
    assert torch.allclose(rm0, rm1, atol=1e-5, rtol=1e-5) and torch.allclose(rv0, rv1, atol=1e-5, rtol=1e-4)
    if dtype == torch.float32:
        # against float64: the fused kernels must be at least as accurate as the two-kernel fp32 path
        xd = x.double().requires_grad_(True)
        yd = F.gelu(F.batch_norm(xd, None, None, w.double(), b.double(), True, 0.1, 1e-5))
        yd.backward(dy.double())
        assert (y0.double() - yd).abs().max() < 1e-5
        assert (dx0.double() - xd.grad).abs().max() < 1e-4


def test_run_modules_fuses_norm_and_gelu():
    """layers.run_modules: conv1x1 -> native norm + GELU == the stack as written (stock modules)."""
    from gkgnet_b200 import layers
    torch.manual_seed(3)
    conv = torch.nn.Conv2d(80, 160, 1).cuda()
    bn = layers.BatchNorm2d(160).cuda()
    act = torch.nn.GELU()
    x = _cl(torch.randn(4, 80, 24, 24, device="cuda"))
    ref_bn = torch.nn.BatchNorm2d(160).cuda()
    ref_bn.load_state_dict(bn.state_dict())
    xa = x.clone().requires_grad_(True)
    ya = layers.run_modules([conv, bn, act], xa)
    ya.sum().backward()
    ga = (xa.grad.clone(), conv.weight.grad.clone(), bn.weight.grad.clone(), bn.bias.grad.clone())
    conv.zero_grad()
    xb = x.clone().requires_grad_(True)
    yb = act(ref_bn(conv(xb)))
    yb.sum().backward()
    assert torch.allclose(ya, yb, atol=1e-4, rtol=1e-4)
    for a, b in zip(ga, (xb.grad, conv.weight.grad, ref_bn.weight.grad, ref_bn.bias.grad)):
        assert torch.allclose(a, b, atol=5e-3, rtol=2e-3)
    assert torch.allclose(bn.running_var, ref_bn.running_var, rtol=1e-5)


def test_sync_batch_norm_single_process_is_batch_norm():
    """layers.SyncBatchNorm (the reference's default norm type) without a process group == layers.BatchNorm2d."""
    from gkgnet_b200 import layers
    torch.manual_seed(2)
    a, b = layers.SyncBatchNorm(160).cuda(), layers.BatchNorm2d(160).cuda()
    b.load_state_dict(a.state_dict())
    x = _cl(torch.randn(4, 160, 20, 20, device="cuda").to(torch.bfloat16))
    gelu = torch.nn.GELU()
    for mods in ((a,), (a, gelu)):
        other = tuple(b if m is a else m for m in mods)
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        ya, yb = layers.run_modules(mods, xa), layers.run_modules(other, xb)
        ya.float().sum().backward(); yb.float().sum().backward()
        assert torch.equal(ya, yb) and torch.equal(xa.grad, xb.grad)
    assert torch.equal(a.running_var, b.running_var) and int(a.num_batches_tracked) == 2
