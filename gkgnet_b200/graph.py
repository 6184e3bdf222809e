"""kNN graph modules -- host-side mirror of vig_model/torch_edge.py backed by
``gkg_knn_graph`` (no distance matrix is ever materialised)."""
from __future__ import annotations

import torch
from torch import nn

from . import ops


def _to_tokens(t):
    """(P, D, N, 1) reference layout -> (P, N, D) token-major (copy unless channels-last)."""
    return t.squeeze(-1).transpose(1, 2)


def edge_index_from_neighbors(nn_idx):
    """int32 (P, N, k) neighbour ids -> reference ``edge_index`` (2, P, N, k) int64 with the
    centre ids underneath (torch_edge.py:85-86, 105-106)."""
    P, N, k = nn_idx.shape
    center = torch.arange(N, device=nn_idx.device).view(1, N, 1).expand(P, N, k)
    return torch.stack((nn_idx.long(), center), dim=0)


class DenseDilated(nn.Module):
    """Keep every ``dilation``-th neighbour (torch_edge.py:126-149)."""

    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation = dilation
        self.stochastic = stochastic
        self.epsilon = epsilon
        self.k = k

    def forward(self, edge_index):
        if self.stochastic and self.training and torch.rand(1) < self.epsilon:
            pick = torch.randperm(self.k * self.dilation)[:self.k]
            return edge_index[:, :, :, pick]
        return edge_index[:, :, :, ::self.dilation]


class DenseDilatedKnnGraph(nn.Module):
    """Dilated kNN graph, same constructor / forward as torch_edge.py:152-176.

    forward(x (P, D, N, 1), y (P, D, M, 1) | None, relative_pos (1, N, M) | None)
    -> edge_index (2, P, N, k) int64.  ``neighbors`` is the compact entry point used by
    the fused modules: token-major input, int32 (B*G, N, k) output, grouping done by the
    kernel.
    """

    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation = dilation
        self.stochastic = stochastic
        self.epsilon = epsilon
        self.k = k
        self._dilated = DenseDilated(k, dilation, stochastic, epsilon)
        self.algo = ops._lib.KNN_AUTO

    def _stochastic_now(self):
        return self.stochastic and self.training and bool(torch.rand(1) < self.epsilon)

    @torch.no_grad()
    def neighbors(self, x, y=None, relative_pos=None, groups=1, separable=None):
        if self._stochastic_now():
            # random k of the k*d nearest (torch_edge.py:141-144): needs the full sorted list
            full = ops.knn_graph(x, y, relative_pos, groups=groups, k=self.k * self.dilation,
                                 dilation=1, algo=self.algo)
            pick = torch.randperm(self.k * self.dilation, device=full.device)[:self.k]
            return full[:, :, pick].contiguous()
        return ops.knn_graph(x, y, relative_pos, groups=groups, k=self.k, dilation=self.dilation,
                             algo=self.algo, separable=separable)

    def forward(self, x, y=None, relative_pos=None):
        nn_idx = self.neighbors(_to_tokens(x), None if y is None else _to_tokens(y), relative_pos)
        return edge_index_from_neighbors(nn_idx)
