"""GPU: the tcgen05 grouped 1x1 FC (+ folded BN + GELU) against the PyTorch ops of the reference's
BasicConv (torch_nn.py:57-81) on the same inputs.  bf16 tolerance 2e-2 (north-star)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference(x, conv, bn, act):
    """fp32 reference of Conv2d(groups=4) -> BatchNorm2d(eval) -> act on token-major x (R, 2C)."""
    y = torch.nn.functional.conv2d(x.float().t().reshape(1, -1, x.shape[0], 1), conv.weight.float(), conv.bias.float(),
                                   groups=4)
    y = torch.nn.functional.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps)
    if act == "gelu":
        y = torch.nn.functional.gelu(y)
    elif act == "relu":
        y = torch.relu(y)
    return y.reshape(-1, x.shape[0]).t()


@pytest.mark.parametrize("C2,rows,act", [(160, 1000, "gelu"), (160, 128 * 37, "gelu"), (320, 777, "gelu"),
                                         (64, 300, "relu"), (160, 5, None),
                                         (800, 1296 * 2 + 5, "gelu"),      # stage 3 of pvig_s: CG = 200, two column passes
                                         (1280, 324 * 3, "gelu"),          # stage 4: CG = 320
                                         (480, 700, "gelu"),               # arch 't' stage 3: CG = 120 (wide kernel, one pass)
                                         (96, 200, "gelu"), (192, 333, None),    # arch 't' stages 1 - 2
                                         # >= 2 tiles per SM: the TMA kernel (3-D tensor map copy-in, bulk-store staging)
                                         (160, 128 * 300 + 77, "gelu"),    # ragged last tile, 2 - 3 tiles per CTA, CG = 40
                                         (160, 128 * 148 * 7 + 1, None),   # 7 - 8 tiles per CTA: every stage reused
                                         (320, 128 * 297 + 5, "gelu"),     # CG = 80: one accumulator set, two stages
                                         (96, 40000, "gelu"),              # CG = 24: K padded 24 -> 32 (reads the next group)
                                         (192, 38000, "relu"), (64, 38001, None)])
def test_grouped_fc_matches_conv_bn_act(C2, rows, act):
    from gkgnet_b200 import ops
    torch.manual_seed(C2 + rows)
    conv = torch.nn.Conv2d(C2, C2, 1, groups=4).cuda()
    bn = torch.nn.BatchNorm2d(C2).cuda().eval()
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.5)
        bn.running_var.uniform_(0.5, 2.0)
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.3)
        conv.bias.normal_(0, 0.3)
    x = torch.randn(rows, C2, device="cuda").to(torch.bfloat16)
    assert ops.grouped_fc_supported(C2)
    scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).float()
    shift = ((conv.bias - bn.running_mean) * scale + bn.bias).float()
    w_op = ops.grouped_fc_weights(conv.weight.detach(), scale.detach())
    got = ops.grouped_fc(x, w_op, shift.detach(), act)
    want = _reference(x, conv, bn, act)
    err = (got.float() - want).abs().max().item()
    assert err < 2e-2 * max(1.0, want.abs().max().item()), err


@pytest.mark.parametrize("C2,rows", [(160, 128 * 300 + 77), (320, 128 * 297 + 5), (96, 40000), (192, 128 * 296),
                                     (256, 128 * 296 + 3)])     # CG = 64: two accumulator sets, two stages, direct stores
def test_grouped_fc_tma_kernel_is_bit_identical_to_the_small_launch_kernels(C2, rows):
    """>= 2 tiles per SM take the TMA kernel (tensor-map copy-in, staged bulk stores); slices of the same rows take the
    cp.async kernels.  Same MMAs in the same K order, same epilogue: the bf16 outputs must be equal bit for bit."""
    from gkgnet_b200 import ops
    torch.manual_seed(C2 * 7 + rows)
    w = torch.randn(C2, C2 // 4, 1, 1, device="cuda") * 0.2
    scale = torch.rand(C2, device="cuda") + 0.5
    shift = torch.randn(C2, device="cuda") * 0.3
    x = (torch.randn(rows, C2, device="cuda") * 2).to(torch.bfloat16)
    w_op = ops.grouped_fc_weights(w, scale)
    for act in ("gelu", None):
        big = ops.grouped_fc(x, w_op, shift, act)
        small = torch.cat([ops.grouped_fc(x[i:i + 4096].contiguous(), w_op, shift, act) for i in range(0, rows, 4096)])
        assert torch.equal(big.view(torch.int16), small.view(torch.int16)), (act, (big.float() - small.float()).abs().max().item())


def test_grouped_fc_supported_widths():
    from gkgnet_b200 import ops
    for c2 in (96, 160, 192, 320, 480, 800, 1280, 1536):   # every MRConv width of the pvig archs
        assert ops.grouped_fc_supported(c2), c2
    assert not ops.grouped_fc_supported(100)       # 4 conv groups of a multiple of 8 channels


def test_mrconv_eval_uses_fused_fc_and_matches_unfused():
    """MRConv2d in eval mode under bf16 autocast takes the fused tensor-core FC; it must agree with the
    unfused conv -> BN -> GELU stack of the same module (which the reference runs)."""
    import gkgnet_b200 as G
    from gkgnet_b200 import ops
    torch.manual_seed(3)
    G.set_norm_type("BN")
    B, C, H, W, groups, k = 2, 80, 24, 24, 2, 9
    m = G.vertex.MRConv2d(C, 2 * C, "gelu", "batch", True).cuda().eval()
    with torch.no_grad():
        m.nn[1].running_mean.normal_(0, 0.3)
        m.nn[1].running_var.uniform_(0.5, 2.0)
        m.nn[0].bias.normal_(0, 0.2)
    xt = torch.randn(B, H * W, C, device="cuda").to(torch.bfloat16)
    idx = ops.knn_graph(xt, None, None, groups=groups, k=k, dilation=1)
    calls = []
    orig = ops.grouped_fc
    ops.grouped_fc = lambda *a, **kw: (calls.append(1), orig(*a, **kw))[1]
    try:
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            fused = m.forward_tokens(xt, idx, None, groups=groups, hw=(H, W))
            agg = ops.mr_aggregate(xt, idx, None, groups=groups)
            plain = m.nn(G.vertex.tokens_to_nchw(agg, H, W))
    finally:
        ops.grouped_fc = orig
    assert calls, "the fused FC was not used"
    err = (fused.float() - plain.float()).abs().max().item()
    assert err < 2e-2 * max(1.0, plain.float().abs().max().item()), err


@pytest.mark.parametrize("C2,R", [(160, 3000), (320, 1000), (800, 1296 + 17), (1280, 324 * 2), (480, 515), (96, 64),
                                  (160, 128 * 300 + 9), (320, 128 * 296 + 64)])    # TMA kernel (forward and data gradient)
def test_grouped_fc_train_gradients_match_conv(C2, R):
    """Training form (forward, data gradient, tcgen05 weight gradient, bias gradient) against Conv2d(groups=4)
    evaluated in fp32 on the same bf16 inputs: 2e-2 of the scale of each tensor (north-star bf16 tolerance)."""
    from gkgnet_b200 import ops
    torch.manual_seed(5)
    conv = torch.nn.Conv2d(C2, C2, 1, groups=4).cuda()
    x = torch.randn(R, C2, device="cuda").to(torch.bfloat16).requires_grad_(True)
    g = torch.randn(R, C2, device="cuda").to(torch.bfloat16)
    out = ops.grouped_fc_train(x, conv.weight, conv.bias)
    out.backward(g)
    gx, gw, gb = x.grad.clone(), conv.weight.grad.clone(), conv.bias.grad.clone()
    x.grad = None; conv.weight.grad = None; conv.bias.grad = None
    xf = x.detach().float().requires_grad_(True)
    wb = conv.weight.detach().to(torch.bfloat16).float().requires_grad_(True)      # the kernel multiplies bf16 weights
    ref = torch.nn.functional.conv2d(xf.t().reshape(1, C2, R, 1), wb, conv.bias.float(), groups=4).reshape(C2, R).t()
    ref.backward(g.float())

    def close(a, b, tol=2e-2):
        return (a.float() - b.float()).abs().max().item() <= tol * max(1.0, b.float().abs().max().item())
    assert close(out, ref), "forward"
    assert close(gx, xf.grad), "data gradient"
    assert close(gw, wb.grad), "weight gradient"
    assert close(gb, g.float().sum(0)), "bias gradient"
