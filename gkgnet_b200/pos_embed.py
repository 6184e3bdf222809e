"""Relative position bias of a Grapher block, computed from its separable structure.

The reference builds ``relative_pos = -bicubic(2 * PE @ PE.T / C, size=(n, n // r^2))``
(pos_embed.py:21-29, torch_vertex.py:309-315) through a dense ``n x n`` float64 product
(3.4 GB at n = 20736).  With the 2-D sin-cos embedding (first half of the channels encodes
the column ``w``, second half the row ``h``; pos_embed.py:38-64)

    PE[q] . PE[k] = F[w_q, w_k] + F[h_q, h_k],   F[i, j] = sum_f cos((i - j) * omega_f)

so only an ``S x S`` table is needed.  The bicubic resize acts on the *flattened* key axis
(the query axis keeps its size, which is the identity for bicubic), i.e. each output key is
a 4-tap combination of flattened source keys; we apply those taps directly.
"""
from __future__ import annotations

import functools
import math

import numpy as np
import torch


def axis_table(channels: int, side: int) -> np.ndarray:
    """F[i, j] = sum_f cos((i - j) * omega_f), float64, (side, side).

    omega_f = 10000^(-f / (C/4)) for f in [0, C/4): get_1d_sincos_pos_embed_from_grid
    (pos_embed.py:67-85) is called with embed_dim = C/2 per axis."""
    nfreq = (channels // 2) // 2
    omega = 1.0 / 10000 ** (np.arange(nfreq, dtype=np.float64) / (channels // 2 / 2.0))
    diff = np.arange(side, dtype=np.float64)[:, None] - np.arange(side, dtype=np.float64)[None, :]
    return np.cos(diff[:, :, None] * omega[None, None, :]).sum(-1)


def _cubic_coeffs(t: torch.Tensor, a: float = -0.75):
    """ATen's bicubic convolution weights (UpSample.h cubic_convolution1/2), fp32."""
    def c1(x):
        return ((a + 2) * x - (a + 3)) * x * x + 1

    def c2(x):
        return ((a * x - 5 * a) * x + 8 * a) * x - 4 * a
    return c2(t + 1), c1(t), c1(1 - t), c2(2 - t)


def relative_pos_table(channels: int, n: int, r: int = 1) -> torch.Tensor:
    """fp32 (1, n, n // r^2): the value the reference stores in ``Grapher.relative_pos``."""
    return _relative_pos_table(int(channels), int(n), int(r)).clone()


@functools.lru_cache(maxsize=8)
def _relative_pos_table(channels: int, n: int, r: int) -> torch.Tensor:
    side = int(n ** 0.5)
    if side * side != n:
        raise ValueError(f"n={n} is not a square grid")
    F = axis_table(channels, side)
    scale = 2.0 / (2 * ((channels // 2) // 2) * 2)       # 2 / PE.shape[1]
    q = np.arange(n)
    hq, wq = q // side, q % side
    m_out = n // (r * r)

    def columns(kk):
        hk, wk = kk // side, kk % side
        full = scale * (F[wq][:, wk] + F[hq][:, hk])                 # float64 (n, len(kk))
        return torch.from_numpy(full.astype(np.float32))

    if m_out == n:
        return -columns(q).unsqueeze(0)
    ratio = torch.tensor(n / m_out, dtype=torch.float32)
    src = ratio * (torch.arange(m_out, dtype=torch.float32) + 0.5) - 0.5
    base = torch.floor(src)
    coeffs = _cubic_coeffs(src - base)
    out = torch.zeros((n, m_out), dtype=torch.float32)
    for tap, w in zip(range(-1, 3), coeffs):
        kk = (base.long() + tap).clamp(0, n - 1).numpy()
        out += columns(kk) * w.unsqueeze(0)
    return -out.unsqueeze(0)
