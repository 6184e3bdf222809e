"""GKGNet backbone -- host-side mirror of mmcls/models/backbones/gkgnet.py.

Same class name, constructor keys, ``forward`` contract and state-dict layout as the
reference (gkgnet.py:121-284), so ``dict(type='GKGNet', choice='s', k=9, ...)`` configs and
pvig/GKGNet checkpoints load unchanged.  Differences, all outside the numerics:
  * no hard-coded ``.cuda()`` (gkgnet.py:264): label ids live in a non-persistent buffer;
  * activations run in ``channels_last`` so the graph kernels see token-major rows;
  * every Grapher / GrapherLabel calls the sm_100a kernels (gkgnet_b200.ops).
"""
from __future__ import annotations

import torch
from torch import nn

from .layers import DropPath, FoldedSequential, act_layer, build_norm_layer, norm_cfg, run_modules
from .registry import BACKBONES, register_into_mmcls
from .vertex import Grapher, GrapherLabel


def _conv_bn(cin, cout, k=1, stride=1, padding=0):
    return [nn.Conv2d(cin, cout, k, stride=stride, padding=padding),
            build_norm_layer(norm_cfg, cout, postfix=1)[1]]


class FFN(nn.Module):
    """1x1-conv feed-forward block with residual (gkgnet.py:46-72)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act="relu", drop_path=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = FoldedSequential(*_conv_bn(in_features, hidden_features))
        self.act = act_layer(act)
        self.fc2 = FoldedSequential(*_conv_bn(hidden_features, out_features))
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def forward(self, x):
        if self.training or torch.is_grad_enabled():
            h = run_modules(list(self.fc1) + [self.act], x)          # norm + GELU of fc1 in one pass
        else:
            h = self.act(self.fc1(x))
        out = self.fc2(h)
        return self.drop_path.add_residual(out, x) if isinstance(self.drop_path, DropPath) else self.drop_path(out) + x


class Stem(nn.Module):
    """Overlapping-conv image embedding, /4 resolution (gkgnet.py:74-101)."""

    def __init__(self, img_size=224, in_dim=3, out_dim=768, act="relu"):
        super().__init__()
        self.convs = FoldedSequential(
            *_conv_bn(in_dim, out_dim // 2, 3, 2, 1), act_layer(act),
            *_conv_bn(out_dim // 2, out_dim, 3, 2, 1), act_layer(act),
            *_conv_bn(out_dim, out_dim, 3, 1, 1))

    def forward(self, x):
        return self.convs(x)


class Downsample(nn.Module):
    """Stride-2 3x3 conv + norm between pyramid stages (gkgnet.py:103-118)."""

    def __init__(self, in_dim=3, out_dim=768):
        super().__init__()
        self.conv = FoldedSequential(*_conv_bn(in_dim, out_dim, 3, 2, 1))

    def forward(self, x):
        return self.conv(x)


_COMMON = dict(k=9, conv="mr", act="gelu", norm="batch", bias=True, dropout=0.0, use_dilation=True,
               epsilon=0.2, use_stochastic=False, blocks=[2, 2, 6, 2], emb_dims=1024)


@BACKBONES.register_module()
class GKGNet(nn.Module):
    arch_settings = {
        "t": dict(_COMMON, channels=[48, 96, 240, 384]),
        "s": dict(_COMMON, channels=[80, 160, 400, 640]),
    }

    def __init__(self, choice="s", k=9, k_label_gcn=9, use_multi_group=True, backbone_multi_group=True,
                 num_group=2, drop_path=0.0, n_classes=1000, out_indices=(3,), size=576, num_gcn=1,
                 pretrain_path=None, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg
        opt = self.arch_settings[choice]
        act, norm, bias = opt["act"], opt["norm"], opt["bias"]
        epsilon, stochastic, conv = opt["epsilon"], opt["use_stochastic"], opt["conv"]
        blocks, channels = opt["blocks"], opt["channels"]
        self.n_blocks = sum(blocks)
        reduce_ratios = [4, 2, 1, 1]
        dpr = [v.item() for v in torch.linspace(0, drop_path, self.n_blocks)]
        max_dilation = 49 // k                                   # gkgnet.py:183

        self.register_buffer("label_input", torch.arange(n_classes).view(1, -1), persistent=False)
        self.label_lt = nn.Embedding(n_classes, channels[0], padding_idx=None)
        ends = [sum(blocks[:i + 1]) + i - 1 for i in range(len(blocks))]
        self.layer_index = ends                                  # [1, 4, 11, 14]
        self.out_indices = [ends[i] for i in out_indices]

        self.stem = Stem(out_dim=channels[0], act=act)
        self.pos_embed = nn.Parameter(torch.zeros(1, channels[0], size // 4, size // 4))
        hw = size // 4 * size // 4

        def label_head(stage, rate):
            return GrapherLabel(channels[stage], k_label_gcn, 1, "mr", act, norm, bias, stochastic,
                                epsilon, reduce_ratios[stage], n=hw, drop_path=rate, relative_pos=False,
                                num_nodes=n_classes, use_multi_group=use_multi_group, num_group=num_group)

        stages, heads, lifts = [], [], []
        idx = 0
        for i, depth in enumerate(blocks):
            if i < len(blocks) - 1:
                heads.append(nn.Sequential(label_head(i, dpr[idx])))
                lifts.append(nn.Sequential(nn.Linear(channels[i], channels[i + 1])))
            else:
                heads.append(nn.ModuleList(label_head(i, dpr[idx]) for _ in range(num_gcn)))
            if i > 0:
                stages.append(Downsample(channels[i - 1], channels[i]))
                hw = hw // 4
            for _ in range(depth):
                stages.append(nn.Sequential(
                    Grapher(channels[i], k, min(idx // 4 + 1, max_dilation), conv, act, norm, bias,
                            stochastic, epsilon, reduce_ratios[i], n=hw, drop_path=dpr[idx],
                            relative_pos=True, use_multi_group=backbone_multi_group, num_group=num_group),
                    FFN(channels[i], channels[i] * 4, act=act, drop_path=dpr[idx])))
                idx += 1
        self.backbone = nn.Sequential(*stages)
        self.gcn_label = nn.Sequential(*heads)
        self.ffn_label = nn.Sequential(*lifts)
        self.gap = nn.AdaptiveAvgPool2d((1, 1))
        self.model_init()

    def model_init(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
                m.weight.requires_grad = True
                if m.bias is not None:
                    m.bias.data.zero_()
                    m.bias.requires_grad = True

    def init_weights(self):
        """Pretrained loading is done by the caller's checkpoint loader (mmcv) via init_cfg."""
        return None

    def forward(self, inputs):
        labels = self.label_lt(self.label_input.expand(inputs.size(0), -1))
        if inputs.is_cuda:
            # channels-last from the first convolution on: cuDNN's NHWC kernels, and the stem's norms take the
            # native channels-last statistics kernels instead of ATen's NCHW ones (5.8 ms of a training step)
            inputs = inputs.contiguous(memory_format=torch.channels_last)
        x = self.stem(inputs) + self.pos_embed
        x = x.contiguous(memory_format=torch.channels_last)
        stage = 0
        edge_index = None
        for i, layer in enumerate(self.backbone):
            x = layer(x)
            if i in self.layer_index:
                for head in self.gcn_label[stage]:
                    labels, edge_index = head(labels, x)
                if stage < 3:
                    labels = self.ffn_label[stage](labels)
                stage += 1
        return labels, torch.flatten(self.gap(x), 1), edge_index


register_into_mmcls(GKGNet)
