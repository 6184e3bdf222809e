"""GPU parity: key pooling (`y = avg_pool2d(x, r, r)` of torch_vertex.py:194-196) through the
C ABI against torch's own avg_pool2d: the fp32 CPU result is the oracle (1e-6 relative: the summation order is
the same, the division is one rounding), and in bf16 the result must equal ATen's CUDA kernel bit for bit
(same fp32 accumulation order, one rounding)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _tokens(x4):
    B, C, H, W = x4.shape
    return x4.permute(0, 2, 3, 1).reshape(B, H * W, C)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,C,H,W,r", [
    (2, 80, 24, 24, 4),       # stage-1 geometry (r = 4)
    (2, 160, 12, 12, 2),      # stage-2 geometry (r = 2)
    (1, 7, 9, 11, 2),         # odd channels (scalar path), H and W not divisible by r (floor mode)
    (3, 20, 10, 7, 3),        # 8-byte vector path for bf16, ragged W
    (1, 16, 4, 4, 4),         # a single window
])
def test_pool_keys_forward_backward(dtype, B, C, H, W, r):
    from gkgnet_b200 import ops
    g = torch.Generator().manual_seed(11)
    x4 = torch.randn(B, C, H, W, generator=g).to(dtype)
    Ho, Wo = H // r, W // r

    xc = x4.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = ops.pool_keys(_tokens(xc), H, W, r)
    assert tuple(y.shape) == (B, Ho * Wo, C) and y.dtype == dtype

    xo = x4.float().requires_grad_(True)
    want = _tokens(F.avg_pool2d(xo, r, r))
    if dtype == torch.float32:
        assert torch.allclose(y.detach().cpu(), want.detach(), rtol=1e-6, atol=1e-7)
    else:
        assert torch.equal(y.detach().cpu(), want.detach().to(dtype))
        xa = x4.cuda().contiguous(memory_format=torch.channels_last)
        assert torch.equal(y.detach(), _tokens(F.avg_pool2d(xa, r, r)))

    w = torch.randn(B, Ho * Wo, C, generator=g).to(dtype)
    y.backward(w.cuda())
    (want * w.float()).sum().backward()
    got = _tokens(xc.grad).float().cpu()
    ref = _tokens(xo.grad)
    if dtype == torch.float32:
        assert torch.allclose(got, ref, rtol=1e-6, atol=1e-7)
    else:
        assert torch.equal(got, ref.to(dtype).float())


def test_pool_keys_rejects_bad_arguments():
    from gkgnet_b200 import ops
    x = torch.randn(1, 16, 8, device="cuda")
    with pytest.raises(ValueError):
        ops.pool_keys(x, 4, 5, 2)            # H*W != nodes
    with pytest.raises(ValueError):
        ops.pool_keys(x, 4, 4, 8)            # window larger than the map
    with pytest.raises(RuntimeError):
        ops.pool_keys(x.cpu(), 4, 4, 2)      # no CPU fallback


def test_pool_keys_full_size_stage1():
    """BASELINE config 2 shape (B=32, C=80, 144x144, r=4): equals ATen's kernel bitwise; constant maps stay
    constant; pooling the gradient of a pooled sum conserves mass."""
    from gkgnet_b200 import ops
    B, C, H, W, r = 32, 80, 144, 144, 4
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, H * W, C, device="cuda", generator=g).bfloat16().requires_grad_(True)
    y = ops.pool_keys(x, H, W, r)
    x4 = x.detach().view(B, H, W, C).permute(0, 3, 1, 2)
    assert torch.equal(y.detach(), _tokens(F.avg_pool2d(x4, r, r)))
    ones = torch.full_like(x, 0.375)
    assert torch.equal(ops.pool_keys(ones, H, W, r), torch.full_like(y, 0.375))
    go = torch.randn_like(y)
    y.backward(go)
    assert abs(float(x.grad.double().sum() - go.double().sum())) < 1e-2 * float(go.double().abs().sum()) ** 0.5 + 8
