"""Small building blocks shared by the graph modules (mirror of the reference's
vig_model/torch_nn.py interface: ``act_layer``, ``norm_layer``, ``BasicConv``,
``batched_index_select``; plus ``DropPath`` which the reference takes from timm)."""
from __future__ import annotations

import torch
from torch import nn

# The reference hard-codes SyncBN (torch_nn.py:8, torch_vertex.py:14, gkgnet.py:23).
# SyncBatchNorm and BatchNorm2d share state-dict keys, so the throughput runs may switch
# to per-GPU statistics with set_norm_type('BN') without touching checkpoints.
norm_cfg = dict(type="SyncBN", requires_grad=True)


def set_norm_type(kind: str):
    if kind not in ("SyncBN", "BN"):
        raise ValueError(kind)
    norm_cfg["type"] = kind


def build_norm_layer(cfg, num_features, postfix=""):
    """Stand-in for mmcv.cnn.build_norm_layer with the two types the reference uses."""
    kind = cfg.get("type", "SyncBN")
    if kind == "SyncBN":
        layer = nn.SyncBatchNorm(num_features)
    elif kind == "BN":
        layer = nn.BatchNorm2d(num_features)
    else:
        raise NotImplementedError(f"norm type [{kind}] is not supported")
    for p in layer.parameters():
        p.requires_grad = cfg.get("requires_grad", True)
    return f"bn{postfix}", layer


def act_layer(act, inplace=False, neg_slope=0.2, n_prelu=1):
    """Activation factory, same names/errors as torch_nn.py:13-29."""
    table = {
        "relu": lambda: nn.ReLU(inplace),
        "leakyrelu": lambda: nn.LeakyReLU(neg_slope, inplace),
        "prelu": lambda: nn.PReLU(num_parameters=n_prelu, init=neg_slope),
        "gelu": lambda: nn.GELU(),
        "hswish": lambda: nn.Hardswish(inplace),
    }
    try:
        return table[act.lower()]()
    except KeyError:
        raise NotImplementedError("activation layer [%s] is not found" % act) from None


def norm_layer(norm, nc):
    """2-D normalisation factory, torch_nn.py:32-42."""
    kind = norm.lower()
    if kind == "batch":
        return build_norm_layer(norm_cfg, nc, postfix=1)[1]
    if kind == "instance":
        return nn.InstanceNorm2d(nc, affine=False)
    raise NotImplementedError("normalization layer [%s] is not found" % norm)


class BasicConv(nn.Sequential):
    """Stack of grouped(4) 1x1 convs, each followed by norm / act / dropout
    (torch_nn.py:57-81).  Sub-module order -- and therefore the state-dict keys
    ``0.weight, 0.bias, 1.<bn>`` -- matches the reference."""

    def __init__(self, channels, act="relu", norm=None, bias=True, drop=0.0):
        mods = []
        for cin, cout in zip(channels[:-1], channels[1:]):
            mods.append(nn.Conv2d(cin, cout, 1, bias=bias, groups=4))
            if norm is not None and norm.lower() != "none":
                mods.append(norm_layer(norm, channels[-1]))
            if act is not None and act.lower() != "none":
                mods.append(act_layer(act))
            if drop > 0:
                mods.append(nn.Dropout2d(drop))
        super().__init__(*mods)
        self.reset_parameters()

    def reset_parameters(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d)) and m.weight is not None:
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)


class DropPath(nn.Module):
    """Stochastic depth per sample (timm.models.layers.DropPath semantics)."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.dim() - 1)
        return x * x.new_empty(shape).bernoulli_(keep).div_(keep)

    def extra_repr(self):
        return f"drop_prob={self.drop_prob}"


def batched_index_select(x, idx):
    """Reference-compatible gather (torch_nn.py:84-105): x (B, C, M, 1), idx (B, N, k)
    -> (B, C, N, k).  Kept for API completeness; the fused aggregate kernel never
    materialises this tensor."""
    B, C, M = x.shape[:3]
    flat = x.reshape(B, C, M)
    N, k = idx.shape[1:]
    out = torch.gather(flat, 2, idx.reshape(B, 1, N * k).expand(B, C, N * k))
    return out.reshape(B, C, N, k)
