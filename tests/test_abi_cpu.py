"""CPU: the C-ABI library loads and exports every symbol include/gkg_abi.h declares; host
logic of the Python mirror (state-dict layout, registry, error behaviour)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gkg_abi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gkg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from gkgnet_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert "gkg_knn_graph" in names and "gkg_mr_aggregate_bwd" in names
    for name in names:
        assert hasattr(lib, name), name
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert _lib.load().gkg_abi_version() == 2


def test_workspace_query_is_host_only():
    from gkgnet_b200 import _lib
    lib = _lib.load()
    small = lib.gkg_knn_workspace_bytes(1, 2, 64, 16, 8, 3, 1, 0, _lib.GKG_F32, _lib.KNN_EXACT_FP32)
    big = lib.gkg_knn_workspace_bytes(32, 2, 20736, 1296, 40, 9, 1, 0, _lib.GKG_F32, _lib.KNN_EXACT_FP32)
    assert 0 < small < big
    # normalised operands + norms for queries and keys
    assert big >= 4 * 64 * (20736 + 1296) * 41


def test_kernels_refuse_cpu_tensors():
    from gkgnet_b200 import ops
    x = torch.randn(1, 8, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.knn_graph(x, None, None, groups=1, k=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.mr_aggregate(x, torch.zeros(1, 8, 2, dtype=torch.int32), None, groups=1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.pool_keys(x, 4, 2, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.grouped_fc(x.bfloat16(), x.bfloat16(), x[0, 0], "gelu")


def test_state_dict_layout_matches_reference():
    import gkgnet_b200 as G
    from tests._util import load_golden
    g = load_golden("gkgnet_s192")
    net = G.build_backbone(dict(type="GKGNet", choice="s", n_classes=7, size=192))
    sd = net.state_dict()
    assert [str(k) for k in g["keys"]] == list(sd.keys())
    want = [tuple(int(s) for s in str(x).split(",") if s) for x in g["shapes"]]
    assert want == [tuple(v.shape) for v in sd.values()]
    rel = [k for k, p in net.named_parameters() if not p.requires_grad]
    assert rel and all(k.endswith("relative_pos") for k in rel)


def test_grapher_dilation_schedule_and_registry():
    import gkgnet_b200 as G
    net = G.GKGNet(choice="t", n_classes=5, size=192, k=18, k_label_gcn=18)
    dil = [m.graph_conv.d for m in net.modules() if isinstance(m, G.Grapher)]
    assert dil == [1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2]      # max_dilation = 49 // 18 = 2
    assert net.layer_index == [1, 4, 11, 14]
    with pytest.raises(KeyError):
        G.build_backbone(dict(type="NoSuchNet"))
    with pytest.raises(NotImplementedError):
        G.GraphConv2d(8, 16, conv="nope")
    with pytest.raises(NotImplementedError):
        G.act_layer("swishh")
    head = G.build_head(dict(type="LabelQueryHead", num_classes=7, in_channels=16))
    lab, gap = torch.randn(2, 7, 16), torch.randn(2, 16)
    from oracle import gkg_oracle as O
    want = O.label_query_score(head.state_dict(), lab, gap)
    assert torch.allclose(head.get_score((lab, gap)), want, atol=1e-6)
    tgt = (torch.rand(2, 7) < 0.3).float()
    got = head.forward_train((lab, gap), tgt)
    ref = O.head_losses(head.state_dict(), lab, gap, tgt)
    for k in ("bce_loss", "asy_loss"):
        assert torch.allclose(got[k], ref[k], atol=1e-5, rtol=1e-5)


def test_basic_conv_folds_its_norm_in_eval():
    """BasicConv (torch_nn.py:57-81) on the library path: the eval-mode fold of the grouped conv + batch norm equals
    the modules as written, and the state-dict keys stay the reference's."""
    from gkgnet_b200.layers import BasicConv
    torch.manual_seed(0)
    m = BasicConv([32, 32], "gelu", "batch", True)
    assert list(m.state_dict().keys())[:3] == ["0.weight", "0.bias", "1.weight"]
    m[1].running_mean.normal_()
    m[1].running_var.uniform_(0.5, 2.0)
    m[1].weight.data.normal_()
    m[1].bias.data.normal_()
    m.eval()
    x = torch.randn(2, 32, 5, 3)
    with torch.no_grad():
        got = m(x)
        want = x
        for mod in m:
            want = mod(want)
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)


def test_folded_sequential_matches_plain_modules_in_eval():
    """Eval-mode Conv2d -> BatchNorm folding (layers.FoldedSequential) is the same function as running the two
    modules, keeps the reference's state-dict keys and stays out of the way in training / under autograd."""
    from torch import nn
    from gkgnet_b200.layers import FoldedSequential
    torch.manual_seed(0)
    seq = FoldedSequential(nn.Conv2d(6, 10, 3, stride=2, padding=1), nn.BatchNorm2d(10), nn.GELU(),
                           nn.Conv2d(10, 4, 1), nn.BatchNorm2d(4))
    with torch.no_grad():
        for m in seq:
            if isinstance(m, nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.5); m.running_var.uniform_(0.5, 2.0)
                m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.3)
    assert list(seq.state_dict().keys())[:2] == ["0.weight", "0.bias"]
    x = torch.randn(2, 6, 9, 9)
    seq.eval()
    with torch.no_grad():
        folded = seq(x)
    plain = nn.Sequential.forward(seq, x)           # grad enabled -> the modules as written
    assert torch.allclose(folded, plain, atol=1e-5, rtol=1e-5)
    seq.train()
    assert seq(x).requires_grad


def test_drop_path_residual_form_equals_the_two_step_form():
    """layers.DropPath.add_residual(x, s) == DropPath(x) + s (timm semantics: per-sample keep mask scaled by 1 / keep),
    same random draws; identity in eval mode and for rate 0."""
    import torch
    from gkgnet_b200.layers import DropPath
    dp = DropPath(0.3).train()
    x, s = torch.randn(64, 5, 3, 3), torch.randn(64, 5, 3, 3)
    torch.manual_seed(7)
    a = dp.add_residual(x, s)
    torch.manual_seed(7)
    b = dp(x) + s
    assert torch.allclose(a, b, atol=1e-6)
    dropped = (a == s).flatten(1).all(1)
    assert 0 < int(dropped.sum()) < 64                      # some samples dropped, some kept
    kept = ~dropped
    assert torch.allclose(a[kept], s[kept] + x[kept] / 0.7, atol=1e-6)
    dp.eval()
    assert torch.equal(dp.add_residual(x, s), x + s)
    assert torch.equal(DropPath(0.0).train().add_residual(x, s), x + s)
    mixed = DropPath(0.3).train().add_residual(x.to(torch.bfloat16), s)      # bf16 branch onto an fp32 residual stream
    assert mixed.dtype == torch.float32


def test_drop_path_residual_gradients():
    """The fused residual form routes gradients like the two-step form: g * mask / keep to the branch (in the branch's
    dtype), g to the shortcut."""
    import torch
    from gkgnet_b200.layers import DropPath
    dp = DropPath(0.4).train()
    x = torch.randn(32, 4, 2, 2, requires_grad=True)
    s = torch.randn(32, 4, 2, 2, requires_grad=True)
    w = torch.randn(32, 4, 2, 2)
    torch.manual_seed(3)
    (dp.add_residual(x, s) * w).sum().backward()
    ga, gs = x.grad.clone(), s.grad.clone()
    x.grad = s.grad = None
    torch.manual_seed(3)
    ((dp(x) + s) * w).sum().backward()
    assert torch.allclose(ga, x.grad, atol=1e-6) and torch.allclose(gs, s.grad, atol=1e-6)
    xb = torch.randn(8, 4, 2, 2).to(torch.bfloat16).requires_grad_(True)
    sb = torch.randn(8, 4, 2, 2, requires_grad=True)
    dp.add_residual(xb, sb).sum().backward()
    assert xb.grad.dtype == torch.bfloat16 and sb.grad.dtype == torch.float32
