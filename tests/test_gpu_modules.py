"""GPU parity at module level: Grapher / GrapherLabel / GKGNet (reference class API,
reference weights from the golden fixtures) against the reference outputs."""
import pytest
import torch

from oracle import gkg_oracle as O
from oracle.gen_golden import det_tensor
from tests._util import load_golden

pytestmark = pytest.mark.gpu


def _same_sets(a, b):
    return (torch.sort(a, -1).values == torch.sort(b, -1).values).all(-1).float().mean().item()


@pytest.mark.parametrize("name,mg", [("grapher_r2", True), ("grapher_r1", True), ("grapher_nogroup", False)])
def test_grapher_eval(name, mg):
    import gkgnet_b200 as G
    g = load_golden(name)
    m = G.Grapher(16, g["k"], g["dilation"], "mr", "gelu", "batch", True, False, 0.2, g["r"], 64, 0.0,
                  True, mg, 2)
    missing = m.load_state_dict(g["sd"], strict=True)
    m = m.cuda().eval()
    for fmt in (torch.contiguous_format, torch.channels_last):
        out = m(g["x"].cuda().contiguous(memory_format=fmt))
        assert torch.allclose(out.cpu(), g["out"], atol=1e-4, rtol=1e-4)
    if "edge_index" in g:
        m.graph_conv.compact_edge_index = False
        _, ei = m.graph_conv(m.fc1(g["x"].cuda()), m.relative_pos)
        assert ei.dtype == torch.int64 and torch.equal(ei.cpu(), g["edge_index"])


def test_grapher_relative_pos_matches_reference_parameter():
    import gkgnet_b200 as G
    g = load_golden("grapher_r2")
    m = G.Grapher(16, 3, 1, "mr", "gelu", "batch", True, False, 0.2, 2, 64, 0.0, True, True, 2)
    assert torch.allclose(m.relative_pos.data, g["sd"]["relative_pos"], atol=1e-6, rtol=0)
    assert not m.relative_pos.requires_grad


@pytest.mark.parametrize("name,mg", [("grapher_label", True), ("grapher_label_nogroup", False)])
def test_grapher_label_eval(name, mg):
    import gkgnet_b200 as G
    g = load_golden(name)
    m = G.GrapherLabel(16, g["k"], 1, "mr", "gelu", "batch", True, False, 0.2, 1, 64, 0.0, False, 5, mg, 2)
    m.load_state_dict(g["sd"], strict=True)
    m = m.cuda().eval()
    out, ei = m(g["labels"].cuda(), g["features"].cuda())
    assert torch.allclose(out.cpu(), g["out"], atol=1e-4, rtol=1e-4)
    assert ei.dtype == torch.int64 and torch.equal(ei.cpu(), g["edge_index"])


def test_grapher_train_grads():
    import gkgnet_b200 as G
    g = load_golden("grapher_train")
    m = G.Grapher(16, g["k"], g["dilation"], "mr", "gelu", "batch", True, False, 0.2, g["r"], 64, 0.0,
                  True, True, 2)
    m.load_state_dict(g["sd"], strict=True)
    m = m.cuda().train()
    x = g["x"].cuda().requires_grad_(True)
    out = m(x)
    assert torch.allclose(out.cpu(), g["out"], atol=1e-4, rtol=1e-4)
    (out * g["w"].cuda()).sum().backward()
    assert torch.allclose(x.grad.cpu(), g["grad_x"], atol=2e-4, rtol=1e-3)
    params = dict(m.named_parameters())
    for k, want in g["grads"].items():
        assert torch.allclose(params[k].grad.cpu(), want, atol=5e-4, rtol=2e-3), k


def _frac_close(a, b, atol=1e-3, rtol=1e-3):
    return ((a - b).abs() <= atol + rtol * b.abs()).float().mean().item()


def test_gkgnet_s192_eval():
    """Whole backbone with the reference's constructor keys and a deterministic state dict.

    GPU and CPU convolutions differ in the last bits, which flips a few near-tie neighbours;
    with random weights such a flip changes that node's features by O(1) and the deviation
    compounds through 16 graph layers.  So (a) every layer is checked on the ORACLE's input
    (teacher forcing): all but a tiny fraction of its outputs must match, and (b) the
    end-to-end outputs must match the reference fixture in shape/dtype and stay strongly
    correlated."""
    import gkgnet_b200 as G
    g = load_golden("gkgnet_s192")
    net = G.build_backbone(dict(type="GKGNet", choice="s", k=9, k_label_gcn=9, drop_path=0.0,
                                n_classes=7, size=192))
    sd = net.state_dict()
    new = {k: (v if k.endswith("relative_pos") else det_tensor(k, v.shape, v.dtype)) for k, v in sd.items()}
    net.load_state_dict(new, strict=True)
    net = net.cuda().eval()
    img = det_tensor("img", (2, 3, 192, 192)) * (3 * 192 * 192) ** 0.5
    trace = {}
    with torch.no_grad():
        want = O.gkgnet_forward({k: v.cpu() for k, v in net.state_dict().items()}, img, "s", 9, 9, 2, trace=trace)
    assert torch.allclose(want[1], g["gap"], atol=2e-4, rtol=1e-3)      # oracle == reference fixture

    sd_cpu = {k: v.cpu() for k, v in net.state_dict().items()}
    plan, _, _ = O.backbone_plan("s", 9)
    with torch.no_grad():
        for i, layer in enumerate(net.backbone):
            xin = trace[f"backbone.{i}.in"]
            out = layer(xin.cuda())
            frac = _frac_close(out.cpu(), trace[f"backbone.{i}.out"])
            assert frac > 0.99, (i, frac)
            if plan[i][0] == "down":
                continue
            # WHY a position differs: the 1x1 convs and the aggregate act per node, so a node's output can only
            # deviate when its neighbour ids do -- and every id deviation must be a distance tie within the band
            # the GPU / CPU rounding of fc1 opens (1e-4 relative), never a wrong neighbour
            _, C, kk, dil, r = plan[i]
            gr, p = layer[0], f"backbone.{i}.0."
            B, _, H, W = xin.shape
            h = gr.fc1(xin.cuda())
            xt = G.vertex.nchw_to_tokens(h)
            yt = G.ops.pool_keys(xt, H, W, r) if r > 1 else None
            ids = G.ops.knn_graph(xt, yt, gr.relative_pos, groups=2, k=kk, dilation=dil).cpu()      # (B*2, N, k)
            h_or = O.conv1x1_bn(sd_cpu, p + "fc1.", xin)
            y_or = torch.nn.functional.avg_pool2d(h_or, r, r) if r > 1 else None
            D, N = C // 2, H * W
            xr = h_or.reshape(B * 2, D, N, 1)
            yr = None if y_or is None else y_or.reshape(B * 2, D, -1, 1)
            dist = O.knn_distance_matrix(xr, yr, sd_cpu[p + "relative_pos"])
            rep = O.check_knn_against_distances(ids, dist, kk, dil, 1e-4)
            assert rep["rows_bad"] == 0, (i, rep)
            ei_or = O.dense_dilated_knn_graph(xr, yr, kk, dil, sd_cpu[p + "relative_pos"])[0]
            node_ids_differ = (ids != ei_or).any(-1).view(B, 2, N).any(1)                             # (B, N)
            o, w = out.cpu(), trace[f"backbone.{i}.out"]
            node_out_differs = ((o - w).abs() > 1e-3 + 1e-3 * w.abs()).any(1).reshape(B, N)
            unexplained = (node_out_differs & ~node_ids_differ).sum().item()
            assert unexplained == 0, (i, unexplained, int(node_out_differs.sum()), int(node_ids_differ.sum()))
        for j in range(4):
            feats = trace[f"backbone.{net.layer_index[j]}.out"].cuda()
            out, ei = net.gcn_label[j][0](trace[f"gcn_label.{j}.0.in"].cuda(), feats)
            frac = _frac_close(out.cpu(), trace[f"gcn_label.{j}.0.out"])
            assert frac > 0.97, (j, frac)
        lab, gap, ei = net(img.cuda())
    assert tuple(lab.shape) == (2, 7, 640) and tuple(gap.shape) == (2, 640) and tuple(ei.shape) == (4, 7, 9)
    assert ei.dtype == torch.int64 and lab.dtype == torch.float32
    cos = torch.nn.functional.cosine_similarity(gap.cpu().flatten(), g["gap"].flatten(), dim=0).item()
    assert cos > 0.98, cos


def test_gkgnet_576_bf16_smoke():
    """BASELINE config 3/4 plumbing at batch 2: GKGNet-576 fwd+bwd under bf16 autocast."""
    import gkgnet_b200 as G
    G.set_norm_type("BN")
    try:
        net = G.GKGNet(choice="s", n_classes=80, size=576, drop_path=0.1).cuda().train()
        head = G.LabelQueryHead(80, 640).cuda()
    finally:
        G.set_norm_type("SyncBN")
    img = torch.randn(2, 3, 576, 576, device="cuda")
    tgt = (torch.rand(2, 80, device="cuda") < 0.04).float()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        feats = net(img)
        losses = head.forward_train(feats, tgt)
    loss = sum(losses.values())
    loss.backward()
    assert torch.isfinite(loss)
    missing = [n for n, p in net.named_parameters() if p.requires_grad and p.grad is None]
    assert not missing, missing[:5]


def test_gkgnet_t_bf16_train_and_eval():
    """Arch 't' (channels 48 / 96 / 240 / 384: grouped-FC widths 2C/4 = 24, 48, 120, 192; group widths D = 24, 48, 120,
    192 -- none of them the 's' widths the kernels were tuned on; ADVICE r1: C2 = 480 used to pass the support check and
    fail at launch): training step and eval forward under bf16 autocast, every parameter gets a gradient, the eval fast
    paths agree with the module stack run with autograd on."""
    import gkgnet_b200 as G
    G.set_norm_type("BN")
    try:
        torch.manual_seed(0)
        net = G.GKGNet(choice="t", n_classes=80, size=192, drop_path=0.1).cuda().train()
        head = G.LabelQueryHead(80, 384).cuda()
    finally:
        G.set_norm_type("SyncBN")
    img = torch.randn(2, 3, 192, 192, device="cuda")
    tgt = (torch.rand(2, 80, device="cuda") < 0.1).float()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = sum(head.forward_train(net(img), tgt).values())
    loss.backward()
    assert torch.isfinite(loss)
    missing = [n for n, p in net.named_parameters() if p.requires_grad and p.grad is None]
    assert not missing, missing[:5]
    net.eval()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        with torch.no_grad():
            fast = net(img)[1].float()                      # folded norms, fused grouped FC
        slow = net(img)[1].float()                          # modules as written (autograd on)
    cos = torch.nn.functional.cosine_similarity(fast.flatten(), slow.flatten(), dim=0).item()
    assert cos > 0.98, cos


@pytest.mark.parametrize("name", ["grapher_r2"])   # (the r = 1 fixture has 8-dim groups: bf16 ties flip too many neighbours)
def test_grapher_eval_bf16_fast_paths_agree_with_module_stack(name):
    """Inference fast paths (Conv -> BN folding, 1x1 convs as GEMMs, tcgen05 grouped FC with folded norm +
    GELU) under no_grad + bf16 autocast against the same module run as written (autograd on) and against the
    reference's fp32 output.  bf16 bar: 2e-2 of the output scale; near-tied neighbours may flip when the
    kNN inputs differ in the last bf16 bit, so a small fraction of positions is allowed to differ more."""
    import gkgnet_b200 as G
    g = load_golden(name)
    m = G.Grapher(16, g["k"], g["dilation"], "mr", "gelu", "batch", True, False, 0.2, g["r"], 64, 0.0, True, True, 2)
    m.load_state_dict(g["sd"], strict=True)
    m = m.cuda().eval()
    x = g["x"].cuda().contiguous(memory_format=torch.channels_last)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        with torch.no_grad():
            fast = m(x).float()
        plain = m(x).float().detach()
    scale = g["out"].abs().max().item()
    for want in (plain.cpu(), g["out"]):
        bad = ((fast.cpu() - want).abs() > 2e-2 * scale).float().mean().item()
        assert bad < 0.02, bad
