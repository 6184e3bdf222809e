"""CPU: pin oracle/gkg_oracle.py against fixtures produced by the unmodified reference
(oracle/gen_golden.py).  These are the only 'golden vectors' that exist for this path:
the reference ships no tests."""
import numpy as np
import pytest
import torch

from oracle import gkg_oracle as O
from oracle.gen_golden import det_tensor
from tests._util import load_golden


def test_knn_xy_matches_reference():
    g = load_golden("knn_xy")
    ei = O.dense_dilated_knn_graph(g["x"], g["y"], g["k"], g["dilation"], g["relative_pos"])
    assert ei.dtype == torch.int64 and tuple(ei.shape) == (2, 4, 64, 3)
    assert torch.equal(ei, g["edge_index"])
    dist = O.knn_distance_matrix(g["x"], g["y"], g["relative_pos"])
    assert torch.allclose(dist, g["dist"], atol=1e-6, rtol=0)
    rep = O.check_knn_against_distances(g["edge_index"][0], dist, g["k"], g["dilation"])
    assert rep["rows_bad"] == 0 and rep["rows_differ"] == 0


def test_knn_self_matches_reference():
    g = load_golden("knn_self")
    ei = O.dense_dilated_knn_graph(g["x"], None, g["k"], g["dilation"], g["relative_pos"])
    assert torch.equal(ei, g["edge_index"])
    ei1 = O.dense_dilated_knn_graph(g["x"], None, g["k"], 1, None)
    assert torch.equal(ei1, g["edge_index_nobias_d1"])
    # without bias every node is its own nearest neighbour
    assert torch.equal(ei1[0, :, :, 0], ei1[1, :, :, 0])


def test_mrconv_matches_reference():
    g = load_golden("mrconv")
    x, y, ei, sd = g["x"], g["y"], g["edge_index"], g["sd"]
    assert torch.equal(O.gather_neighbors(y, ei[0]), g["x_j"])
    agg = O.mr_aggregate(x, ei, y, in_channels=16)
    assert tuple(agg.shape) == (2, 32, 64, 1)
    # odd channels carry the max-relative term, even channels the centre feature
    assert torch.equal(agg[:, 1::2].reshape(4, 8, 64, 1), g["maxrel"])
    assert torch.equal(agg[:, 0::2].reshape(4, 8, 64, 1), x)
    out = O.mr_conv2d(sd, "", x, ei, y, 16)
    assert torch.allclose(out, g["out"], atol=1e-6, rtol=1e-6)
    ei_self = O.dense_dilated_knn_graph(x, None, 3, 1, None)
    assert torch.allclose(O.mr_conv2d(sd, "", x, ei_self, None, 16), g["out_self"], atol=1e-6, rtol=1e-6)
    # max-relative identity used by the CUDA kernel: max_j(x_j - x_i) == max_j(x_j) - x_i bitwise
    alt = g["x_j"].max(-1, keepdim=True).values - x
    assert torch.equal(alt, g["maxrel"])


@pytest.mark.parametrize("name", ["grapher_r2", "grapher_r1"])
def test_grapher_matches_reference(name):
    g = load_golden(name)
    out = O.grapher(g["sd"], "", g["x"], g["k"], g["dilation"], g["r"], 2, True)
    assert torch.allclose(out, g["out"], atol=1e-5, rtol=1e-5)
    h = O.conv1x1_bn(g["sd"], "fc1.", g["x"])
    _, ei = O.dygraph_conv(g["sd"], "graph_conv.", h, g["sd"]["relative_pos"], g["k"], g["dilation"],
                           g["r"], 2)
    assert torch.equal(ei, g["edge_index"])


def test_grapher_nogroup_matches_reference():
    g = load_golden("grapher_nogroup")
    out = O.grapher(g["sd"], "", g["x"], g["k"], g["dilation"], g["r"], 2, False)
    assert torch.allclose(out, g["out"], atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("name,mg", [("grapher_label", True), ("grapher_label_nogroup", False)])
def test_grapher_label_matches_reference(name, mg):
    g = load_golden(name)
    out, ei = O.grapher_label(g["sd"], "", g["labels"], g["features"], g["k"], 2, mg)
    assert torch.allclose(out, g["out"], atol=1e-5, rtol=1e-5)
    assert torch.equal(ei, g["edge_index"])


def test_relative_pos_matches_reference():
    g = load_golden("relative_pos")
    for key, want in g.items():
        if not key.startswith("rel_"):
            continue
        c, n, r = (int(s[1:]) for s in key.split("_")[1:])
        got = O.relative_pos_table(c, n, r)
        assert got.shape == want.shape
        assert torch.allclose(got, want, atol=1e-6, rtol=0), key


def test_grapher_train_grads_match_reference():
    g = load_golden("grapher_train")
    sd = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k
                                      and "relative_pos" not in k) for k, v in g["sd"].items()}
    x = g["x"].clone().requires_grad_(True)
    out = O.grapher(sd, "", x, g["k"], g["dilation"], g["r"], 2, True, training=True)
    assert torch.allclose(out, g["out"], atol=1e-5, rtol=1e-5)
    (out * g["w"]).sum().backward()
    assert torch.allclose(x.grad, g["grad_x"], atol=1e-4, rtol=1e-4)
    for k, want in g["grads"].items():
        assert torch.allclose(sd[k].grad, want, atol=2e-4, rtol=1e-3), k


def test_gkgnet_s192_matches_reference():
    g = load_golden("gkgnet_s192")
    sd = {}
    for key, shp in zip(g["keys"], g["shapes"]):
        key = str(key)
        shape = tuple(int(s) for s in str(shp).split(",") if s)
        if key.endswith("relative_pos"):
            continue
        sd[key] = det_tensor(key, shape)
    plan, _, channels = O.backbone_plan("s", 9)
    side = g["size"] // 4
    for i, item in enumerate(plan):
        if item[0] == "down":
            side //= 2
        else:
            sd[f"backbone.{i}.0.relative_pos"] = O.relative_pos_table(item[1], side * side, item[4])
    img = det_tensor("img", (2, 3, 192, 192)) * (3 * 192 * 192) ** 0.5
    with torch.no_grad():
        lab, gap, ei = O.gkgnet_forward(sd, img, "s", 9, 9, 2)
    assert torch.allclose(gap, g["gap"], atol=2e-4, rtol=1e-3)
    assert torch.allclose(lab, g["label_emb"], atol=2e-4, rtol=1e-3)
    assert ei.shape == g["edge_index"].shape
    same = (torch.sort(ei, -1).values == torch.sort(g["edge_index"], -1).values).all(-1).float().mean()
    assert same > 0.99


@pytest.mark.parametrize("conv", ["edge", "sage", "gin", "gat"])
def test_graphconv_variants_match_reference(conv):
    """The graph convolutions GKGNet does not instantiate (torch_vertex.py:16-150), un-grouped as the reference
    builds them: oracle restatement against the reference's own output, separate keys and self keys."""
    g = load_golden("gconv_" + conv)
    fn = {"edge": O.edge_conv2d, "sage": O.graph_sage, "gin": O.gin_conv2d, "gat": O.graph_atten}[conv]
    out = fn(g["sd"], "gconv.", g["x"], g["edge_index"], g["y"])
    assert torch.allclose(out, g["out"], atol=1e-5, rtol=1e-5)
    out = fn(g["sd"], "gconv.", g["x"], g["edge_index_self"], None)
    assert torch.allclose(out, g["out_self"], atol=1e-5, rtol=1e-5)
