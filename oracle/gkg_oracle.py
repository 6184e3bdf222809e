"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the GKGNet graph hot path.

A functional, state-dict driven restatement (plain PyTorch, fp32, CPU) of the
algorithm implemented by the reference files under
``mmcls/models/backbones/vig_model/`` and ``mmcls/models/backbones/gkgnet.py``.
Every function cites the reference file:line it follows.  Nothing here is imported by
the product package ``gkgnet_b200``: only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it, and only as
the checker / CPU baseline.

Pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the *reference itself*, executed in the authoring
container through ``oracle/ref_shim.py`` and committed as ``tests/golden/*.npz`` by
``oracle/gen_golden.py`` (see tests/test_oracle_golden.py).

Conventions (reference layout): node features are ``(B, C, N, 1)`` NCHW tensors,
``edge_index`` is ``(2, B*G, N, k)`` int64 with ``[0]`` = neighbour ids, ``[1]`` =
centre ids.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# kNN graph construction                                   (reference: torch_edge.py)
# --------------------------------------------------------------------------------------


def l2_normalize(x, dim=1, eps=1e-12):
    """``F.normalize(x, p=2, dim=1)`` as used at torch_edge.py:167-168,173."""
    denom = x.norm(p=2, dim=dim, keepdim=True).clamp_min(eps)
    return x / denom


def sq_distance(xn, yn):
    """Squared distance matrix in the reference's association order.

    ``(x_sq + (-2 x y^T)) + y_sq^T`` -- torch_edge.py:48-51 (xy), :18-20 (self),
    :31-36 (chunked).  xn: (P, N, D), yn: (P, M, D) -> (P, N, M).
    """
    inner = -2 * torch.matmul(xn, yn.transpose(2, 1))
    x_sq = torch.sum(xn * xn, dim=-1, keepdim=True)
    y_sq = torch.sum(yn * yn, dim=-1, keepdim=True)
    return x_sq + inner + y_sq.transpose(2, 1)


def knn_distance_matrix(x, y=None, relative_pos=None):
    """Biased distance matrix the reference ranks on.

    x: (P, D, N, 1); y: (P, D, M, 1) or None (keys = queries); relative_pos (1, N, M).
    Follows DenseDilatedKnnGraph.forward (torch_edge.py:164-176) up to the topk call:
    normalise, transpose to (P, N, D), distance, ``dist += relative_pos``
    (torch_edge.py:79-82, :100-103).
    """
    xn = l2_normalize(x, dim=1).transpose(2, 1).squeeze(-1)
    yn = xn if y is None else l2_normalize(y, dim=1).transpose(2, 1).squeeze(-1)
    dist = sq_distance(xn, yn)
    if relative_pos is not None:
        dist = dist + relative_pos
    return dist


def dense_dilated_knn_graph(x, y=None, k=9, dilation=1, relative_pos=None):
    """edge_index (2, P, N, k) int64 -- torch_edge.py:164-176 + :139-149 + :83-86,104-106.

    k*dilation nearest keys sorted by ascending distance (``topk(-dist)``), centre ids
    stacked underneath, then every ``dilation``-th rank kept.
    """
    with torch.no_grad():
        dist = knn_distance_matrix(x, y, relative_pos)
        P, N, _ = dist.shape
        kd = k * dilation
        nn_idx = torch.topk(-dist, k=kd).indices
        center = torch.arange(N, device=dist.device).view(1, N, 1).expand(P, N, kd)
        edge_index = torch.stack((nn_idx, center), dim=0)
        return edge_index[:, :, :, ::dilation]


# --------------------------------------------------------------------------------------
# gather + max-relative aggregation                 (reference: torch_nn.py, torch_vertex.py)
# --------------------------------------------------------------------------------------


def gather_neighbors(x, idx):
    """out[p, c, n, j] = x[p, c, idx[p, n, j]] -- batched_index_select, torch_nn.py:84-105."""
    P, C, M = x.shape[:3]
    _, N, K = idx.shape
    flat = x.squeeze(-1).transpose(1, 2).reshape(P * M, C)
    rows = (idx + torch.arange(P, device=idx.device).view(P, 1, 1) * M).reshape(-1)
    return flat[rows].view(P, N, K, C).permute(0, 3, 1, 2)


def mr_aggregate(x, edge_index, y=None, in_channels=None):
    """Max-relative features, channel-interleaved -- MRConv2d.forward, torch_vertex.py:47-61.

    x: (P, D, N, 1); edge_index (2, P, N, k); y: (P, D, M, 1) or None.
    Returns (P*D/in_channels, 2*in_channels, N, 1) with channels [x_0, m_0, x_1, m_1, ...]
    where m = max_j (x_j - x_i).
    """
    src = x if y is None else y
    x_i = gather_neighbors(x, edge_index[1])
    x_j = gather_neighbors(src, edge_index[0])
    m = torch.max(x_j - x_i, dim=-1, keepdim=True).values
    in_channels = in_channels or x.shape[1]
    P, D, N, _ = x.shape
    b = P * D // in_channels
    xg = x.reshape(b, in_channels, N, 1)
    mg = m.reshape(b, in_channels, N, 1)
    return torch.stack((xg, mg), dim=2).reshape(b, 2 * in_channels, N, 1)


def batch_norm(sd, prefix, x, training=False, eps=1e-5):
    """(Sync)BatchNorm as produced by ``build_norm_layer`` (torch_nn.py:37); without a
    process group SyncBatchNorm computes plain batch statistics."""
    return F.batch_norm(x, sd[prefix + "running_mean"].clone(), sd[prefix + "running_var"].clone(),
                        sd[prefix + "weight"], sd[prefix + "bias"], training, 0.1, eps)


def conv1x1_bn(sd, prefix, x, training=False, groups=1):
    """``Sequential(Conv2d(.., 1), norm)`` -- e.g. Grapher.fc1/fc2, torch_vertex.py:290-306."""
    x = F.conv2d(x, sd[prefix + "0.weight"], sd[prefix + "0.bias"], groups=groups)
    return batch_norm(sd, prefix + "1.", x, training)


def basic_conv(sd, prefix, x, training=False):
    """BasicConv([2C, 2C], 'gelu', 'batch'): grouped(4) 1x1 conv -> BN -> GELU,
    torch_nn.py:57-70."""
    return F.gelu(conv1x1_bn(sd, prefix, x, training, groups=4))


def mr_conv2d(sd, prefix, x, edge_index, y, in_channels, training=False):
    """MRConv2d.forward incl. ``self.nn`` -- torch_vertex.py:47-62."""
    return basic_conv(sd, prefix + "nn.", mr_aggregate(x, edge_index, y, in_channels), training)


# --------------------------------------------------------------------------------------
# dynamic graph conv wrappers                              (reference: torch_vertex.py)
# --------------------------------------------------------------------------------------


def edge_conv2d(sd, prefix, x, edge_index, y=None, training=False):
    """EdgeConv2d.forward (torch_vertex.py:91-101): max_k nn(cat[x_i, x_j - x_i])."""
    x_i = gather_neighbors(x, edge_index[1])
    x_j = gather_neighbors(x if y is None else y, edge_index[0])
    h = basic_conv(sd, prefix + "nn.", torch.cat([x_i, x_j - x_i], dim=1), training)
    return h.max(-1, keepdim=True).values


def graph_sage(sd, prefix, x, edge_index, y=None, training=False):
    """GraphSAGE.forward (torch_vertex.py:125-131): nn2(cat[x, max_k nn1(x_j)])."""
    x_j = gather_neighbors(x if y is None else y, edge_index[0])
    m = basic_conv(sd, prefix + "nn1.", x_j, training).max(-1, keepdim=True).values
    return basic_conv(sd, prefix + "nn2.", torch.cat([x, m], dim=1), training)


def gin_conv2d(sd, prefix, x, edge_index, y=None, training=False):
    """GINConv2d.forward (torch_vertex.py:143-149): nn((1 + eps) x + sum_k x_j)."""
    x_j = gather_neighbors(x if y is None else y, edge_index[0]).sum(-1, keepdim=True)
    return basic_conv(sd, prefix + "nn.", (1 + sd[prefix + "eps"]) * x + x_j, training)


def graph_atten(sd, prefix, x, edge_index, y=None, training=False):
    """GraphAtten.forward (torch_vertex.py:26-37): softmax_k(a(cat[x_i, x_j])) weighted neighbour mean, interleaved
    with the centre features, then nn."""
    x_i = gather_neighbors(x, edge_index[1])
    x_j = gather_neighbors(x if y is None else y, edge_index[0])
    e = F.conv2d(torch.cat([x_i, x_j], dim=1), sd[prefix + "a.weight"], sd[prefix + "a.bias"]).squeeze(1)
    att = torch.softmax(e, -1)
    m = (att.unsqueeze(-1) * x_j.permute(0, 2, 3, 1)).sum(2).transpose(1, 2).unsqueeze(-1)
    b, c, n, _ = x.shape
    h = torch.cat([x.unsqueeze(2), m.unsqueeze(2)], dim=2).reshape(b, 2 * c, n, 1)
    return basic_conv(sd, prefix + "nn.", h, training)


def dygraph_conv(sd, prefix, x, relative_pos, k, dilation, r, num_group=1, training=False):
    """DyGraphConv2dMultiGroup.forward (torch_vertex.py:191-205); ``num_group=1``
    reproduces DyGraphConv2d.forward (:218-228).  x: (B, C, H, W) -> ((B, 2C, H, W), edge_index)."""
    B, C, H, W = x.shape
    y = None
    if r > 1:
        y = F.avg_pool2d(x, r, r).reshape(B, C, -1, 1)
    xf = x.reshape(B, C, -1, 1)
    D = C // num_group
    xg = xf.reshape(B * num_group, D, -1, 1)
    yg = None if y is None else y.reshape(B * num_group, D, -1, 1)
    edge_index = dense_dilated_knn_graph(xg, yg, k, dilation, relative_pos)
    out = mr_conv2d(sd, prefix + "gconv.", xg, edge_index, yg, C, training)
    return out.reshape(B, -1, H, W), edge_index


def dygraph_label(sd, prefix, x, feats, k, num_group=1, multi_group=True, training=False):
    """DyGraphLabelMultiGroup.forward (torch_vertex.py:266-275) / DyGraphLabel.forward
    (:243-251).  x: (B, C, nodes, 1), feats: (B, C, HW) -> ((B, 2C, nodes, 1), edge_index)."""
    B, C, N, _ = x.shape
    G = num_group if multi_group else 1
    D = C // G
    yg = feats.reshape(B * G, D, -1, 1)
    xg = x.reshape(B * G, D, -1, 1)
    edge_index = dense_dilated_knn_graph(xg, yg, k, 1, None)
    out = mr_conv2d(sd, prefix + "gconv.", xg, edge_index, yg, C, training)
    out = out.reshape(B, 2 * C, -1, 1)
    return out, (edge_index[0] if multi_group else edge_index)


def grapher(sd, prefix, x, k, dilation, r, num_group=2, multi_group=True, training=False):
    """Grapher.forward (torch_vertex.py:325-333); drop_path is identity (eval / p=0).
    ``relative_pos`` comes from the state dict (parameter built at :309-315)."""
    shortcut = x
    x = conv1x1_bn(sd, prefix + "fc1.", x, training)
    rel = sd.get(prefix + "relative_pos")
    x, _ = dygraph_conv(sd, prefix + "graph_conv.", x, rel, k, dilation, r,
                        num_group if multi_group else 1, training)
    x = conv1x1_bn(sd, prefix + "fc2.", x, training)
    return x + shortcut


def ffn(sd, prefix, x, training=False):
    """FFN.forward (gkgnet.py:66-72) / FFNLabel body (torch_vertex.py:352-357)."""
    h = F.gelu(conv1x1_bn(sd, prefix + "fc1.", x, training))
    return conv1x1_bn(sd, prefix + "fc2.", h, training) + x


def grapher_label(sd, prefix, x, features, k, num_group=2, multi_group=True, training=False):
    """GrapherLabel.forward (torch_vertex.py:392-403).  x: (B, nodes, C),
    features: (B, C, H, W) -> (x (B, nodes, C), edge_index)."""
    B, C, H, W = features.shape
    feats = features.reshape(B, C, -1)
    x = x.transpose(2, 1).unsqueeze(-1)
    shortcut = x
    x = conv1x1_bn(sd, prefix + "fc1.", x, training)
    x, edge_index = dygraph_label(sd, prefix + "graph_conv.", x, feats, k, num_group,
                                  multi_group, training)
    x = conv1x1_bn(sd, prefix + "fc2.", x, training) + shortcut
    x = ffn(sd, prefix + "ffn.", x, training)
    return x.transpose(2, 1).squeeze(-1), edge_index


# --------------------------------------------------------------------------------------
# relative position bias                                     (reference: pos_embed.py)
# --------------------------------------------------------------------------------------


def sincos_1d(dim, pos):
    """get_1d_sincos_pos_embed_from_grid, pos_embed.py:67-85 (float64)."""
    omega = np.arange(dim // 2, dtype=np.float64)
    omega /= dim / 2.0
    omega = 1.0 / 10000 ** omega
    out = np.einsum("m,d->md", pos.reshape(-1).astype(np.float64), omega)
    return np.concatenate([np.sin(out), np.cos(out)], axis=1)


def sincos_2d(dim, grid_size):
    """get_2d_sincos_pos_embed, pos_embed.py:38-64: first half encodes the w grid."""
    gh = np.arange(grid_size, dtype=np.float32)
    gw = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape(2, 1, grid_size, grid_size)
    return np.concatenate([sincos_1d(dim // 2, grid[0]), sincos_1d(dim // 2, grid[1])], axis=1)


def relative_pos_table(channels, n, r):
    """The ``relative_pos`` parameter of a Grapher: ``-interp(2 PE PE^T / C)``,
    pos_embed.py:21-29 + torch_vertex.py:309-315.  Returns (1, n, n // r^2) fp32."""
    pe = sincos_2d(channels, int(n ** 0.5))
    rel = np.float32(2 * np.matmul(pe, pe.T) / pe.shape[1])
    t = torch.from_numpy(rel).unsqueeze(0).unsqueeze(1)
    t = F.interpolate(t, size=(n, n // (r * r)), mode="bicubic", align_corners=False)
    return -t.squeeze(1)


# --------------------------------------------------------------------------------------
# whole backbone + head                          (reference: gkgnet.py, label_query_head.py)
# --------------------------------------------------------------------------------------

ARCH = {"t": dict(blocks=[2, 2, 6, 2], channels=[48, 96, 240, 384]),
        "s": dict(blocks=[2, 2, 6, 2], channels=[80, 160, 400, 640])}  # gkgnet.py:122-149


def backbone_plan(choice="s", k=9):
    """Layer sequence of ``GKGNet.backbone`` with per-Grapher (k, dilation, r) --
    gkgnet.py:176-183, 228-239."""
    blocks, channels = ARCH[choice]["blocks"], ARCH[choice]["channels"]
    ratios = [4, 2, 1, 1]
    max_dil = 49 // k
    plan, idx = [], 0
    for i, nb in enumerate(blocks):
        if i > 0:
            plan.append(("down", channels[i - 1], channels[i]))
        for _ in range(nb):
            plan.append(("block", channels[i], k, min(idx // 4 + 1, max_dil), ratios[i]))
            idx += 1
    layer_index = [sum(blocks[:i + 1]) + i - 1 for i in range(len(blocks))]  # gkgnet.py:189
    return plan, layer_index, channels


def gkgnet_forward(sd, img, choice="s", k=9, k_label_gcn=9, num_group=2, num_gcn=1,
                   training=False, trace=None):
    """GKGNet.forward (gkgnet.py:263-284) in eval / drop_path=0 semantics.
    Returns (label_emb (B, n_cls, C4), gap (B, C4), edge_index (B*G, n_cls, k)).
    ``trace`` (a dict) receives the input/output of every backbone layer and label head so a
    test can compare layer by layer without accumulating near-tie neighbour flips."""
    plan, layer_index, channels = backbone_plan(choice, k)
    B = img.shape[0]
    labels = sd["label_lt.weight"].unsqueeze(0).expand(B, -1, -1)
    x = img
    for ci, bi in ((0, 1), (3, 4), (6, 7)):                       # Stem, gkgnet.py:83-101
        stride = 2 if ci < 6 else 1
        x = F.conv2d(x, sd[f"stem.convs.{ci}.weight"], sd[f"stem.convs.{ci}.bias"],
                     stride=stride, padding=1)
        x = batch_norm(sd, f"stem.convs.{bi}.", x, training)
        if ci < 6:
            x = F.gelu(x)
    x = x + sd["pos_embed"]
    j = 0
    edge_index = None
    for i, item in enumerate(plan):
        p = f"backbone.{i}."
        if trace is not None:
            trace[f"backbone.{i}.in"] = x
        if item[0] == "down":                                      # Downsample, gkgnet.py:107-118
            x = F.conv2d(x, sd[p + "conv.0.weight"], sd[p + "conv.0.bias"], stride=2, padding=1)
            x = batch_norm(sd, p + "conv.1.", x, training)
        else:
            _, C, kk, dil, r = item
            x = grapher(sd, p + "0.", x, kk, dil, r, num_group, True, training)
            x = ffn(sd, p + "1.", x, training)
        if trace is not None:
            trace[f"backbone.{i}.out"] = x
        if i in layer_index:                                       # gkgnet.py:272-277
            for g in range(num_gcn if j == 3 else 1):
                if trace is not None:
                    trace[f"gcn_label.{j}.{g}.in"] = labels
                labels, edge_index = grapher_label(sd, f"gcn_label.{j}.{g}.", labels, x,
                                                   k_label_gcn, num_group, True, training)
                if trace is not None:
                    trace[f"gcn_label.{j}.{g}.out"] = labels
            if j < 3:
                labels = F.linear(labels, sd[f"ffn_label.{j}.0.weight"], sd[f"ffn_label.{j}.0.bias"])
            j += 1
    gap = torch.flatten(F.adaptive_avg_pool2d(x, (1, 1)), 1)
    return labels, gap, edge_index


def label_query_score(sd, label_emb, gap, prefix=""):
    """LabelQueryHead.get_score (label_query_head.py:49-57): diagonal of fc1(L) + fc2(gap)."""
    out1 = F.linear(label_emb, sd[prefix + "fc1.weight"], sd[prefix + "fc1.bias"])
    diag = torch.diagonal(out1, dim1=1, dim2=2)
    return diag + F.linear(gap, sd[prefix + "fc2.weight"], sd[prefix + "fc2.bias"])


def asymmetric_loss(score, target, gamma_pos=0.0, gamma_neg=2.0, clip=0.05, eps=1e-8):
    """asymmetric_loss (losses/asymmetric_loss.py:9-72) summed / batch (cls_head.py:44-49)."""
    p = torch.sigmoid(score)
    t = target.type_as(score)
    pt = (1 - p + clip).clamp(max=1) * (1 - t) + p * t
    w = (1 - pt).pow(gamma_pos * t + gamma_neg * (1 - t))
    return (-torch.log(pt.clamp(min=eps)) * w).sum() / score.shape[0]


def label_smooth_bce(score, target, smooth=0.1):
    """LabelSmoothLoss(mode='multi_label') -> BCE-with-logits, summed / batch
    (label_smooth_loss.py:122-126,168-175)."""
    t = target.type_as(score) * (1 - 2 * smooth) + smooth
    return F.binary_cross_entropy_with_logits(score, t, reduction="sum") / score.shape[0]


def head_losses(sd, label_emb, gap, target, prefix=""):
    """LabelQueryHead.forward_train with double_loss (label_query_head.py:70-85)."""
    s = label_query_score(sd, label_emb, gap, prefix)
    return {"bce_loss": label_smooth_bce(s, target), "asy_loss": 10.0 * asymmetric_loss(s, target)}


# --------------------------------------------------------------------------------------
# comparison helpers used by the parity tests
# --------------------------------------------------------------------------------------


def check_knn_against_distances(nn_idx, dist, k, dilation, rtol=1e-6):
    """Validate a kNN result against the oracle's biased distance matrix.

    nn_idx: (P, N, k) integer neighbour ids (after dilation); dist: (P, N, M) oracle
    distances.  The contract (BASELINE.json north_star): index sets bit-identical to the
    reference's except at distance ties within ``rtol`` relative.  Distances are sums of
    O(1) terms (unit vectors, bias in [-1, 0]), so "relative" is taken against
    max(1, |d|).  Returns a dict with the number of rows that differ from the oracle's
    own top-k and the number of rows violating the tie band (must be 0).
    """
    P, N, M = dist.shape
    kd = k * dilation
    nn_idx = torch.as_tensor(nn_idx).long()
    assert nn_idx.shape == (P, N, k), (nn_idx.shape, (P, N, k))
    assert int(nn_idx.min()) >= 0 and int(nn_idx.max()) < M
    d64 = dist.double()
    ref_sorted, ref_idx = torch.sort(d64, dim=-1, stable=True)
    ref_pick = ref_idx[..., :kd:dilation]
    got_d = torch.gather(d64, 2, nn_idx)
    want_d = ref_sorted[..., :kd:dilation]
    tol = rtol * torch.maximum(want_d.abs(), torch.ones_like(want_d))
    # rank-wise distance agreement covers both membership and order (dilation picks ranks)
    bad = (got_d - want_d).abs() > tol
    # neighbour ids must be distinct within a row
    srt = torch.sort(nn_idx, dim=-1).values
    dup = (srt[..., 1:] == srt[..., :-1]).any(-1)
    differ = (nn_idx != ref_pick).any(-1)
    return {"rows": P * N, "rows_differ": int(differ.sum()), "rows_bad": int((bad.any(-1) | dup).sum()),
            "max_excess": float(((got_d - want_d).abs() - tol).clamp(min=0).max())}
