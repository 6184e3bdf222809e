"""Graph convolution modules -- host-side mirror of vig_model/torch_vertex.py.

Class names, constructor arguments, attribute names (hence state-dict keys) and forward
signatures follow the reference so that configs and checkpoints drop in; the work is done
by the CUDA kernels behind :mod:`gkgnet_b200.ops`.  Internally activations are handled
token-major ``(B, N, C)`` (= ``channels_last`` NCHW), which makes the reference's
transpose/contiguous copies (torch_nn.py:102-104) and the ``(B*G, D, N, 1)`` regrouping
copies (torch_vertex.py:199-202) disappear: grouping is an index computation in the kernel.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .graph import DenseDilatedKnnGraph, edge_index_from_neighbors
from .layers import BasicConv, DropPath, FoldedSequential, act_layer, build_norm_layer, norm_cfg, run_modules
from .pos_embed import relative_pos_table


def nchw_to_tokens(x):
    """(B, C, H, W) -> (B, H*W, C) with channel stride 1 (a view for channels_last input)."""
    B, C, H, W = x.shape
    return x.permute(0, 2, 3, 1).reshape(B, H * W, C)


def tokens_to_nchw(t, H, W):
    """(B, H*W, C) token-major -> (B, C, H, W) view in channels_last memory format."""
    B, _, C = t.shape
    return t.view(B, H, W, C).permute(0, 3, 1, 2)


class MRConv2d(nn.Module):
    """Max-Relative graph convolution (reference: torch_vertex.py:38-62).

    ``forward(x, edge_index, y=None)`` accepts the reference layout -- x (P, D, N, 1),
    edge_index (2, P, N, k), y (P, D, M, 1) with P = B*G -- and returns
    (B, 2*in_channels, N, 1).  ``forward_tokens`` is the copy-free entry used by the
    DyGraph* wrappers."""

    def __init__(self, in_channels, out_channels, act="relu", norm=None, bias=True):
        super().__init__()
        self.in_channels = in_channels
        self.nn = BasicConv([in_channels * 2, out_channels], act, norm, bias)

    def forward_tokens(self, xt, nn_idx, yt=None, groups=1, hw=None):
        """xt (B, N, C), nn_idx int32 (B*G, N, k), yt (B, M, C) | None -> (B, 2C', H, W)
        where (H, W) = hw or (N, 1)."""
        agg = ops.mr_aggregate(xt, nn_idx, yt, groups=groups)          # (B, N, 2C) interleaved
        H, W = hw if hw is not None else (xt.shape[1], 1)
        fused = self._fused_fc(agg)
        if fused is not None:
            return tokens_to_nchw(fused, H, W)
        conv = self.nn[0] if len(self.nn) else None
        if (torch.is_grad_enabled() and agg.dtype == torch.bfloat16 and isinstance(conv, nn.Conv2d) and conv.groups == 4
                and conv.kernel_size == (1, 1) and conv.in_channels == conv.out_channels == agg.shape[-1]
                and ops.grouped_fc_supported(conv.out_channels)):
            # training: the FC and its data gradient on the tensor-core kernel, norm / act stay modules
            h = tokens_to_nchw(ops.grouped_fc_train(agg, conv.weight, conv.bias), H, W)
            return run_modules(list(self.nn)[1:], h)                 # norm + GELU in one pass
        return self.nn(tokens_to_nchw(agg, H, W))

    def _fused_fc(self, agg):
        """Eval-mode ``self.nn`` (grouped 1x1 conv -> batch norm -> activation, torch_nn.py:57-81) as ONE
        tensor-core kernel with the norm folded into a per-channel affine map; None when the stack or the
        call does not qualify (training, gradients, fp32 activations, widths beyond the kernel)."""
        if self.training or torch.is_grad_enabled() or agg.dtype != torch.bfloat16 or not agg.is_cuda:
            return None
        mods = list(self.nn)
        if not mods or not isinstance(mods[0], nn.Conv2d) or mods[0].groups != 4 or mods[0].kernel_size != (1, 1):
            return None
        conv, rest = mods[0], mods[1:]
        bn = rest.pop(0) if rest and isinstance(rest[0], (nn.BatchNorm2d, nn.SyncBatchNorm)) else None
        act = None
        if rest and isinstance(rest[0], nn.GELU) and getattr(rest[0], "approximate", "none") == "none":
            act, rest = "gelu", rest[1:]
        elif rest and isinstance(rest[0], nn.ReLU):
            act, rest = "relu", rest[1:]
        if rest or conv.in_channels != conv.out_channels or conv.in_channels != agg.shape[-1]:
            return None
        if bn is not None and (bn.running_mean is None or not bn.track_running_stats):
            return None
        if not ops.grouped_fc_supported(conv.out_channels):
            return None
        tensors = [conv.weight, conv.bias] + ([bn.weight, bn.bias, bn.running_mean, bn.running_var] if bn is not None else [])
        key = tuple((t.data_ptr(), t._version) for t in tensors if t is not None)
        cache = getattr(self, "_fc_cache", None)
        if cache is None or cache[0] != key:
            with torch.no_grad():
                c2 = conv.out_channels
                scale = torch.ones(c2, device=agg.device)
                shift = torch.zeros(c2, device=agg.device) if conv.bias is None else conv.bias.detach().float().clone()
                if bn is not None:
                    g = bn.weight.detach().float() if bn.weight is not None else torch.ones_like(scale)
                    b = bn.bias.detach().float() if bn.bias is not None else torch.zeros_like(scale)
                    scale = g / torch.sqrt(bn.running_var.float() + bn.eps)
                    shift = (shift - bn.running_mean.float()) * scale + b
                cache = (key, ops.grouped_fc_weights(conv.weight.detach(), scale), shift.contiguous())
            self._fc_cache = cache
        return ops.grouped_fc(agg, cache[1], cache[2], act)

    def forward(self, x, edge_index, y=None):
        P, D, N, _ = x.shape
        if self.in_channels % D or (P * D) % self.in_channels:
            raise ValueError(f"cannot regroup {tuple(x.shape)} into {self.in_channels} channels")
        G = self.in_channels // D
        B = P // G

        def regroup(t):  # (B*G, D, n, 1) -> (B, n, G*D)
            n = t.shape[2]
            return t.reshape(B, G, D, n).permute(0, 3, 1, 2).reshape(B, n, G * D)

        nn_idx = edge_index[0].to(torch.int32)
        return self.forward_tokens(regroup(x), nn_idx, None if y is None else regroup(y), groups=G)


class _GatherConvBase(nn.Module):
    """Shared plumbing of the graph convolutions that work on the gathered neighbour rows.  ``forward`` takes the
    reference layout -- x (B, C, N, 1), edge_index (2, B, N, k), y (B, C, M, 1) | None -- ``forward_tokens`` the
    token-major one.  Like the reference's classes they only make sense un-grouped (their ``nn`` is built for the
    full channel count: torch_vertex.py:88-101 leaves the regrouping commented out)."""

    def forward(self, x, edge_index, y=None):
        B, C, N, _ = x.shape
        xt = x.reshape(B, C, N).transpose(1, 2)
        yt = None if y is None else y.reshape(B, C, -1).transpose(1, 2)
        return self.forward_tokens(xt, edge_index[0].to(torch.int32), yt)

    @staticmethod
    def _nchw(t):            # (B, N, k, C) -> (B, C, N, k) view (channels-last memory: no copy)
        return t.permute(0, 3, 1, 2)

    def _check_groups(self, groups):
        if groups != 1:
            raise ValueError(f"{type(self).__name__} is defined for a single channel group (got {groups})")


class EdgeConv2d(_GatherConvBase):
    """max_j nn(cat[x_i, x_j - x_i]) (torch_vertex.py:82-101)."""

    def __init__(self, in_channels, out_channels, act="relu", norm=None, bias=True):
        super().__init__()
        self.in_channels = in_channels
        self.nn = BasicConv([in_channels * 2, out_channels], act, norm, bias)

    def forward_tokens(self, xt, nn_idx, yt=None, groups=1, hw=None):
        self._check_groups(groups)
        x_j = ops.gather_neighbors(xt if yt is None else yt, nn_idx)            # (B, N, k, C)
        x_i = xt.unsqueeze(2).expand_as(x_j)
        h = self.nn(self._nchw(torch.cat((x_i, x_j - x_i), dim=-1)))
        out = h.max(dim=-1, keepdim=True).values                                # (B, C', N, 1)
        return out if hw is None else out.reshape(out.shape[0], out.shape[1], *hw)


class GraphSAGE(_GatherConvBase):
    """nn2(cat[x, max_j nn1(x_j)]) (torch_vertex.py:116-131)."""

    def __init__(self, in_channels, out_channels, act="relu", norm=None, bias=True):
        super().__init__()
        self.nn1 = BasicConv([in_channels, in_channels], act, norm, bias)
        self.nn2 = BasicConv([in_channels * 2, out_channels], act, norm, bias)

    def forward_tokens(self, xt, nn_idx, yt=None, groups=1, hw=None):
        self._check_groups(groups)
        x_j = ops.gather_neighbors(xt if yt is None else yt, nn_idx)
        m = self.nn1(self._nchw(x_j)).max(dim=-1, keepdim=True).values          # (B, C, N, 1)
        x = xt.transpose(1, 2).unsqueeze(-1)
        out = self.nn2(torch.cat((x, m), dim=1))
        return out if hw is None else out.reshape(out.shape[0], out.shape[1], *hw)


class GINConv2d(_GatherConvBase):
    """nn((1 + eps) x + sum_j x_j) (torch_vertex.py:134-150)."""

    def __init__(self, in_channels, out_channels, act="relu", norm=None, bias=True):
        super().__init__()
        self.nn = BasicConv([in_channels, out_channels], act, norm, bias)
        self.eps = nn.Parameter(torch.Tensor([0.0]))

    def forward_tokens(self, xt, nn_idx, yt=None, groups=1, hw=None):
        self._check_groups(groups)
        s = ops.sum_neighbors(xt if yt is None else yt, nn_idx)                 # (B, N, C), no gathered tensor
        h = ((1 + self.eps) * xt + s).transpose(1, 2).unsqueeze(-1)
        out = self.nn(h)
        return out if hw is None else out.reshape(out.shape[0], out.shape[1], *hw)


class GraphAtten(_GatherConvBase):
    """Attention-weighted neighbour mean interleaved with the centre features (torch_vertex.py:16-37)."""

    def __init__(self, in_channels, out_channels, act="relu", norm=None, bias=True, alpha=0.1):
        super().__init__()
        self.in_channels = in_channels
        self.leakyrelu = nn.LeakyReLU(alpha)
        self.nn = BasicConv([in_channels * 2, out_channels], act, norm, bias)
        self.a = nn.Conv2d(in_channels * 2, 1, 1, bias=bias)

    def forward_tokens(self, xt, nn_idx, yt=None, groups=1, hw=None):
        self._check_groups(groups)
        x_j = ops.gather_neighbors(xt if yt is None else yt, nn_idx)            # (B, N, k, C)
        x_i = xt.unsqueeze(2).expand_as(x_j)
        e = self.a(self._nchw(torch.cat((x_i, x_j), dim=-1)))                   # (B, 1, N, k)
        att = torch.softmax(e.reshape(e.shape[0], e.shape[2], e.shape[3]), dim=-1)
        m = (att.unsqueeze(-1) * x_j).sum(2)                                    # (B, N, C)
        B, N, C = xt.shape
        h = torch.stack((xt, m), dim=-1).reshape(B, N, 2 * C)                   # channels [x_0, m_0, x_1, m_1, ...]
        out = self.nn(h.transpose(1, 2).unsqueeze(-1))
        return out if hw is None else out.reshape(out.shape[0], out.shape[1], *hw)


class GraphConv2d(nn.Module):
    """Static graph convolution selector (torch_vertex.py:153-173).  GKGNet fixes ``conv='mr'``
    (gkgnet.py:124,138,205,220): the fused max-relative kernels; the other variants run on the native neighbour
    gather / sum (csrc/gather.cu) with their 1x1 convolutions as library ops."""

    def __init__(self, in_channels, out_channels, conv="edge", act="relu", norm=None, bias=True):
        super().__init__()
        if conv == "edge":
            self.gconv = EdgeConv2d(in_channels, out_channels, act, norm, bias)
        elif conv == "gat":
            self.gconv = GraphAtten(in_channels, out_channels, act, norm, bias)
        elif conv == "mr":
            self.gconv = MRConv2d(in_channels, out_channels, act, norm, bias)
        elif conv == "sage":
            self.gconv = GraphSAGE(in_channels, out_channels, act, norm, bias)
        elif conv == "gin":
            self.gconv = GINConv2d(in_channels, out_channels, act, norm, bias)
        else:
            raise NotImplementedError("conv:{} is not supported".format(conv))

    def forward(self, x, edge_index, y=None):
        return self.gconv(x, edge_index, y)


class _DynamicGraphBase(GraphConv2d):
    #: when True, ``forward`` returns the compact int32 neighbour tensor (B*G, N, k) instead
    #: of the reference's int64 ``edge_index`` (2, B*G, N, k) -- 191 MB per stage-1 layer at
    #: B=32 that Grapher throws away immediately (torch_vertex.py:330).
    compact_edge_index = False

    def _edge_index(self, nn_idx):
        return nn_idx if self.compact_edge_index else edge_index_from_neighbors(nn_idx)


class DyGraphConv2dMultiGroup(_DynamicGraphBase):
    """Dynamic grouped graph convolution over image patches (torch_vertex.py:175-205)."""

    def __init__(self, in_channels, out_channels, kernel_size=9, dilation=1, conv="edge", act="relu",
                 norm=None, bias=True, stochastic=False, epsilon=0.0, r=1, num_head=2):
        super().__init__(in_channels, out_channels, conv, act, norm, bias)
        self.k = kernel_size
        self.d = dilation
        self.r = r
        self.num_head = num_head
        self.dilated_knn_graph = DenseDilatedKnnGraph(kernel_size, dilation, stochastic, epsilon)

    def forward(self, x, relative_pos=None, separable=None):
        B, C, H, W = x.shape
        xt = nchw_to_tokens(x)
        yt = None
        if self.r > 1:
            yt = ops.pool_keys(xt, H, W, self.r)      # avg_pool2d(x, r, r), token-major
        nn_idx = self.dilated_knn_graph.neighbors(xt, yt, relative_pos, groups=self.num_head,
                                                  separable=separable)
        out = self.gconv.forward_tokens(xt, nn_idx, yt, groups=self.num_head, hw=(H, W))
        return out, self._edge_index(nn_idx)


class DyGraphConv2d(DyGraphConv2dMultiGroup):
    """Single-group variant (torch_vertex.py:206-228)."""

    def __init__(self, in_channels, out_channels, kernel_size=9, dilation=1, conv="edge", act="relu",
                 norm=None, bias=True, stochastic=False, epsilon=0.0, r=1):
        super().__init__(in_channels, out_channels, kernel_size, dilation, conv, act, norm, bias,
                         stochastic, epsilon, r, num_head=1)
        del self.num_head
        self.num_head = 1


class DyGraphLabelMultiGroup(_DynamicGraphBase):
    """Label-node <-> patch grouped graph convolution (torch_vertex.py:253-275): queries are
    the label embeddings, keys all patches, dilation 1, no positional bias."""

    def __init__(self, in_channels, out_channels, kernel_size=9, dilation=1, conv="edge", act="relu",
                 norm=None, bias=True, stochastic=False, epsilon=0.0, r=1, num_head=2, bit_graph=True):
        super().__init__(in_channels, out_channels, conv, act, norm, bias)
        self.k = kernel_size
        self.d = dilation
        self.r = r
        self.num_head = num_head
        self.out_channels = out_channels
        self.dilated_knn_graph = DenseDilatedKnnGraph(kernel_size, dilation, stochastic, epsilon)
        self._multi_group_return = True

    def forward_tokens(self, xt, yt):
        nn_idx = self.dilated_knn_graph.neighbors(xt, yt, None, groups=self.num_head)
        out = self.gconv.forward_tokens(xt, nn_idx, yt, groups=self.num_head)
        if self._multi_group_return:
            edge = nn_idx.long()                          # == edge_index[0], torch_vertex.py:275
        else:
            edge = edge_index_from_neighbors(nn_idx)      # DyGraphLabel returns the pair, :251
        return out, edge

    def forward(self, x, y=None):
        B, C, N, _ = x.shape
        xt = x.reshape(B, C, N).transpose(1, 2)
        yt = None if y is None else y.reshape(B, C, -1).transpose(1, 2)
        return self.forward_tokens(xt, yt)


class DyGraphLabel(DyGraphLabelMultiGroup):
    """Single-group variant (torch_vertex.py:229-251); returns the full edge_index."""

    def __init__(self, in_channels, out_channels, kernel_size=9, dilation=1, conv="edge", act="relu",
                 norm=None, bias=True, stochastic=False, epsilon=0.0, r=1):
        super().__init__(in_channels, out_channels, kernel_size, dilation, conv, act, norm, bias,
                         stochastic, epsilon, r, num_head=1)
        self._multi_group_return = False


def _conv_bn(cin, cout):
    return FoldedSequential(nn.Conv2d(cin, cout, 1, stride=1, padding=0),
                            build_norm_layer(norm_cfg, cout, postfix=1)[1])


class Grapher(nn.Module):
    """Grapher block: fc1 -> dynamic graph conv -> fc2 -> residual (torch_vertex.py:278-333)."""

    def __init__(self, in_channels, kernel_size=9, dilation=1, conv="edge", act="relu", norm=None,
                 bias=True, stochastic=False, epsilon=0.0, r=1, n=196, drop_path=0.0,
                 relative_pos=False, use_multi_group=False, num_group=2):
        super().__init__()
        self.channels = in_channels
        self.n = n
        self.r = r
        self.fc1 = _conv_bn(in_channels, in_channels)
        if use_multi_group:
            self.graph_conv = DyGraphConv2dMultiGroup(in_channels, in_channels * 2, kernel_size, dilation,
                                                      conv, act, norm, bias, stochastic, epsilon, r,
                                                      num_head=num_group)
        else:
            self.graph_conv = DyGraphConv2d(in_channels, in_channels * 2, kernel_size, dilation, conv,
                                            act, norm, bias, stochastic, epsilon, r)
        self.graph_conv.compact_edge_index = True
        self.fc2 = _conv_bn(in_channels * 2, in_channels)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.relative_pos = None
        self._sep_cache = None
        if relative_pos:
            self.relative_pos = nn.Parameter(relative_pos_table(in_channels, n, r), requires_grad=False)

    def _get_relative_pos(self, relative_pos, H, W):
        if relative_pos is None or H * W == self.n:
            return relative_pos
        N = H * W
        return F.interpolate(relative_pos.unsqueeze(0), size=(N, N // (self.r * self.r)),
                             mode="bicubic").squeeze(0)

    def _separable(self, relative_pos):
        """Separable factorisation of the bias, fitted once per parameter version (the table is
        a frozen constant; a loaded checkpoint that does not factorise falls back to dense)."""
        if relative_pos is None or relative_pos is not self.relative_pos or not relative_pos.is_cuda:
            return None
        key = (relative_pos.data_ptr(), relative_pos._version, relative_pos.device)
        if self._sep_cache is None or self._sep_cache[0] != key:
            self._sep_cache = (key, ops.fit_separable_bias(relative_pos.detach()))
        return self._sep_cache[1]

    def forward(self, x):
        shortcut = x
        x = self.fc1(x)
        B, C, H, W = x.shape
        rel = self._get_relative_pos(self.relative_pos, H, W)
        x, _ = self.graph_conv(x, rel, self._separable(rel))
        x = self.fc2(x)
        return self.drop_path.add_residual(x, shortcut) if isinstance(self.drop_path, DropPath) else self.drop_path(x) + shortcut


class FFNLabel(nn.Module):
    """Feed-forward block of the label branch (torch_vertex.py:334-360); returns (B, nodes, C)."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act="relu", drop_path=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = _conv_bn(in_features, hidden_features)
        self.act = act_layer(act)
        self.fc2 = _conv_bn(hidden_features, out_features)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def forward(self, x, y=None):
        out = self.fc2(self.act(self.fc1(x)))
        out = self.drop_path(out) + x
        return out.transpose(2, 1).squeeze(-1)


class GrapherLabel(nn.Module):
    """Label-query Grapher: the group-kNN label head (torch_vertex.py:361-403).

    forward(x (B, nodes, C), features (B, C, H, W)) -> (x (B, nodes, C), edge_index)."""

    def __init__(self, in_channels, kernel_size=9, dilation=1, conv="edge", act="relu", norm=None,
                 bias=True, stochastic=False, epsilon=0.0, r=1, n=196, drop_path=0.0,
                 relative_pos=False, num_nodes=80, use_multi_group=False, num_group=2):
        super().__init__()
        self.channels = in_channels
        self.fc1 = _conv_bn(in_channels, in_channels)
        self.fc2 = _conv_bn(in_channels * 2, in_channels)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        if not use_multi_group:
            self.graph_conv = DyGraphLabel(in_channels, in_channels * 2, kernel_size, dilation, conv,
                                           act, norm, bias, stochastic, epsilon, r)
        else:
            self.graph_conv = DyGraphLabelMultiGroup(in_channels, in_channels * 2, kernel_size, dilation,
                                                     conv, act, norm, bias, stochastic, epsilon, r,
                                                     num_head=num_group)
        self.ffn = FFNLabel(in_channels, in_channels * 4, act=act, drop_path=drop_path)

    def forward(self, x, features):
        feats = nchw_to_tokens(features)                       # (B, HW, C), no NCHW round trip
        x = x.transpose(2, 1).unsqueeze(-1)                    # (B, C, nodes, 1)
        shortcut = x
        x = self.fc1(x)
        B, C, N, _ = x.shape
        x, edge_index = self.graph_conv.forward_tokens(x.permute(0, 2, 3, 1).reshape(B, N, C), feats)
        x = self.fc2(x)
        x = self.drop_path.add_residual(x, shortcut) if isinstance(self.drop_path, DropPath) else self.drop_path(x) + shortcut
        return self.ffn(x), edge_index
