"""Functional entry points of the graph hot path, backed by libgkg_b200.so.

Tensors are token-major: features are ``(B, N, C)`` with the channel stride 1 (a
``channels_last`` NCHW tensor viewed as ``(B, H*W, C)`` is exactly that, no copy).
Channel group g of G owns channels ``[g*D, (g+1)*D)``; "problem" ``p = b*G + g`` as in the
reference's ``(B*G, D, N, 1)`` regrouping (torch_vertex.py:197-202).
"""
from __future__ import annotations

import functools

import torch

from . import _lib

_DT = {torch.float32: _lib.GKG_F32, torch.bfloat16: _lib.GKG_BF16}


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _guard(fn):
    """Run ``fn`` with the device of its first CUDA tensor argument current: the library launches on the current
    device (``<<<>>>``, ``cudaFuncSetAttribute``), which need not be the tensors' device in a multi-GPU process."""
    @functools.wraps(fn)
    def wrapper(*args, **kw):
        dev = next((a.device for a in args if isinstance(a, torch.Tensor) and a.is_cuda), None)
        if dev is None:
            return fn(*args, **kw)
        with torch.cuda.device(dev):
            return fn(*args, **kw)
    return wrapper


def _token_major(t: torch.Tensor) -> torch.Tensor:
    """(B, N, C) with stride(C) == 1 and 16-byte aligned rows; copies only if needed."""
    assert t.dim() == 3
    if t.stride(2) != 1 and t.shape[2] != 1:
        t = t.contiguous()
    return t


def _common_dtype(x, y):
    """Mixed precision (e.g. bf16 label queries vs fp32 patch features under autocast): promote
    to the wider type, like the reference's elementwise ops do."""
    if y is not None and y.dtype != x.dtype:
        dt = torch.promote_types(x.dtype, y.dtype)
        x, y = x.to(dt), y.to(dt)
    return x, y


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("gkgnet_b200 kernels need CUDA tensors (no CPU fallback)")


def fit_separable_bias(relative_pos, tol=5e-7):
    """Try to write ``relative_pos[n, m] = A[n % W][m % Kw] + B[n // W][m // Kw]``.

    The reference's table (2-D sin-cos embedding resized along the flattened key axis,
    pos_embed.py:21-29 + torch_vertex.py:309-315) always has this form with ``W = sqrt(N)``
    and ``Kw = W * M / N``; it lets the tensor-core kernel add the bias from registers
    instead of streaming the dense (N, M) matrix for every problem.  Returns
    ``(A, B, W, Kw)`` (fp32, on the table's device) or None when the table does not fit
    within ``tol`` -- the dense matrix is then read (any stored parameter is honoured)."""
    rel = relative_pos.reshape(relative_pos.shape[-2], relative_pos.shape[-1]).float()
    N, M = rel.shape
    W = int(round(N ** 0.5))
    if W * W != N or (W * M) % N:
        return None
    kw = W * M // N
    if kw < 1 or M % kw or M // kw < 1:
        return None
    mh = M // kw
    r4 = rel.view(W, W, mh, kw)                       # [h_q, w_q, m_h, m_w]
    a = r4[0, :, 0, :].contiguous()                   # (W, Kw)
    b = (r4[:, 0, :, 0] - r4[0, 0, 0, 0]).contiguous()  # (W, Mh)
    resid = (r4 - (a.view(1, W, 1, kw) + b.view(W, 1, mh, 1))).abs().max().item()
    if not resid <= tol:
        return None
    return a, b, W, kw


def knn_graph(x, y=None, relative_pos=None, *, groups=1, k=9, dilation=1, algo=_lib.KNN_AUTO,
              separable=None, debug=None):
    """Dilated group-kNN neighbour ids, int32 ``(B*G, N, k)``.

    x: (B, N, C) queries; y: (B, M, C) keys or None (self); relative_pos: fp32 (N, M) /
    (1, N, M) bias or None.  Equivalent to ``DenseDilatedKnnGraph(k, dilation)(x, y,
    relative_pos)[0]`` of the reference (torch_edge.py:164-176) on the regrouped tensors.

    debug (tests only): dict with optional keys ``flags`` (1 = exact re-rank of every row, 3 = every
    row through the brute-force fix-up kernel), ``dist`` (fp32 (B*G, N, M) CUDA tensor receiving the
    approximate distances), ``skip`` / ``ga`` (threshold-sweep overrides); the counters of the call
    come back in ``debug["stats"]``.  Routed through ``gkg_knn_graph_debug``.
    """
    _require_cuda(x, y, relative_pos)
    lib = _lib.load()
    x, y = _common_dtype(x, y)
    x = _token_major(x)
    B, N, C = x.shape
    if C % groups:
        raise ValueError(f"channels {C} not divisible by groups {groups}")
    D = C // groups
    if x.dtype not in _DT:
        raise TypeError(f"unsupported dtype {x.dtype}")
    if y is not None:
        y = _token_major(y)
        if y.dtype != x.dtype or y.shape[0] != B or y.shape[2] != C:
            raise ValueError("keys must match queries in batch, channels and dtype")
        if y.device != x.device:
            raise ValueError("keys and queries must live on the same device")
        M = y.shape[1]
    else:
        M = N
    rel_ptr = None
    if relative_pos is not None:
        rel = relative_pos.reshape(relative_pos.shape[-2], relative_pos.shape[-1])
        if rel.shape != (N, M):
            raise ValueError(f"relative_pos {tuple(rel.shape)} != ({N}, {M})")
        if rel.device != x.device:
            raise ValueError("relative_pos must live on the device of the features")
        rel = rel.to(torch.float32).contiguous()
        rel_ptr = rel.data_ptr()
    sep_a = sep_b = None
    sep_w = sep_kw = 0
    if separable is not None and relative_pos is not None:
        sa, sb, sep_w, sep_kw = separable
        if sa.shape != (sep_w, sep_kw) or sb.shape != (N // sep_w, M // sep_kw):
            raise ValueError("separable bias tables do not match the problem shape")
        sep_a, sep_b = sa.data_ptr(), sb.data_ptr()
    idx = torch.empty((B * groups, N, k), dtype=torch.int32, device=x.device)
    if B * N == 0:
        return idx
    with torch.cuda.device(x.device):
        ws_bytes = lib.gkg_knn_workspace_bytes(B, groups, N, M, D, k, dilation, int(y is None), _DT[x.dtype], algo)
        ws = _workspace(x.device, ws_bytes)
        args = (x.data_ptr(), x.stride(0), x.stride(1),
                y.data_ptr() if y is not None else None,
                y.stride(0) if y is not None else 0, y.stride(1) if y is not None else 0,
                rel_ptr, sep_a, sep_b, sep_w, sep_kw, idx.data_ptr(), B, groups, N, M, D, k, dilation,
                _DT[x.dtype], algo, ws.data_ptr(), ws_bytes, _stream(x))
        if debug is None:
            _lib.check(lib.gkg_knn_graph(*args), "gkg_knn_graph")
        else:
            import ctypes
            stats = (ctypes.c_uint * 3)()
            dist = debug.get("dist")
            rc = lib.gkg_knn_graph_debug(*args, int(debug.get("flags", 0)), None if dist is None else dist.data_ptr(),
                                         stats, int(debug.get("skip", -1)), int(debug.get("ga", 0)))
            _lib.check(rc, "gkg_knn_graph_debug")
            import struct
            debug["stats"] = {"fixups": stats[0], "ambiguous": stats[1],
                              "max_err": struct.unpack("f", struct.pack("I", stats[2]))[0]}
    return idx


# kNN scratch (hundreds of MB at stage 1) is cached per (device, stream): the calls of one stream are ordered, so
# consecutive layers can share one buffer instead of going through the allocator every call.
_WS_CACHE = {}


def _workspace(device, nbytes):
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _WS_CACHE.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _WS_CACHE.pop(key, None)
        buf = torch.empty(int(nbytes * 1.25) if nbytes > (1 << 20) else nbytes, dtype=torch.uint8, device=device)
        _WS_CACHE[key] = buf
    return buf


class _MRAggregate(torch.autograd.Function):
    @staticmethod
    @_guard
    def forward(ctx, x, y, idx, groups):
        lib = _lib.load()
        x = _token_major(x)
        B, N, C = x.shape
        D = C // groups
        k = idx.shape[-1]
        self_keys = y is None
        if not self_keys:
            y = _token_major(y)
        M = N if self_keys else y.shape[1]
        out = torch.empty((B, N, 2 * C), dtype=x.dtype, device=x.device)
        need_grad = any(ctx.needs_input_grad[:2])
        amax = torch.empty((B, N, C), dtype=torch.uint8, device=x.device) if need_grad else None
        rc = lib.gkg_mr_aggregate_fwd(
            x.data_ptr(), x.stride(0), x.stride(1),
            None if self_keys else y.data_ptr(),
            0 if self_keys else y.stride(0), 0 if self_keys else y.stride(1),
            idx.data_ptr(), out.data_ptr(), amax.data_ptr() if amax is not None else None,
            B, groups, N, M, D, k, _DT[x.dtype], _stream(x))
        _lib.check(rc, "gkg_mr_aggregate_fwd")
        ctx.save_for_backward(idx, amax)
        ctx.meta = (B, groups, N, M, D, k, self_keys, x.dtype)
        return out

    @staticmethod
    @_guard
    def backward(ctx, grad_out):
        idx, amax_plane = ctx.saved_tensors
        B, G, N, M, D, k, self_keys, dtype = ctx.meta
        lib = _lib.load()
        C = G * D
        grad_out = grad_out.contiguous()
        gx = torch.empty((B, N, C), dtype=dtype, device=grad_out.device)
        if deterministic_aggregate():
            # 64-bit fixed-point accumulation: integer adds commute, the result is bitwise reproducible
            amax = grad_out.detach().abs().max().float().clamp_min(1e-30)
            scale = torch.exp2(40.0 - torch.ceil(torch.log2(amax))).reshape(1).contiguous()
            gfix = torch.zeros((B, M, C), dtype=torch.int64, device=grad_out.device)
            rc = lib.gkg_mr_aggregate_bwd_det(grad_out.data_ptr(), idx.data_ptr(), amax_plane.data_ptr(), gx.data_ptr(),
                                              gfix.data_ptr(), scale.data_ptr(), B, G, N, M, D, k, _DT[dtype],
                                              _stream(grad_out))
            _lib.check(rc, "gkg_mr_aggregate_bwd_det")
            gy = torch.empty((B, M, C), dtype=torch.float32, device=grad_out.device)
            _lib.check(lib.gkg_fixed_to_float(gfix.data_ptr(), scale.data_ptr(), gy.data_ptr(), gfix.numel(),
                                              _stream(grad_out)), "gkg_fixed_to_float")
        else:
            gy = torch.zeros((B, M, C), dtype=torch.float32, device=grad_out.device)
            rc = lib.gkg_mr_aggregate_bwd(grad_out.data_ptr(), idx.data_ptr(), amax_plane.data_ptr(),
                                          gx.data_ptr(), gy.data_ptr(), B, G, N, M, D, k, _DT[dtype],
                                          _stream(grad_out))
            _lib.check(rc, "gkg_mr_aggregate_bwd")
        if self_keys:
            return (gx.float() + gy).to(dtype), None, None, None
        return gx, gy.to(dtype), None, None


_DETERMINISTIC = None


def set_deterministic_aggregate(flag):
    """True / False: force the deterministic (64-bit fixed-point) scatter of the aggregate backward on / off;
    None (default): follow ``torch.are_deterministic_algorithms_enabled()``."""
    global _DETERMINISTIC
    _DETERMINISTIC = flag


def deterministic_aggregate():
    return torch.are_deterministic_algorithms_enabled() if _DETERMINISTIC is None else bool(_DETERMINISTIC)


def mr_aggregate(x, idx, y=None, *, groups=1):
    """Max-relative aggregation, ``(B, N, 2C)`` with channels ``[x_0, m_0, x_1, m_1, ...]``.

    x: (B, N, C); idx: int32 (B*G, N, k) from :func:`knn_graph`; y: (B, M, C) or None.
    Equivalent to MRConv2d.forward up to (not including) ``self.nn``
    (torch_vertex.py:49-61).  Differentiable w.r.t. x and y.
    """
    _require_cuda(x, y, idx)
    x, y = _common_dtype(x, y)
    if idx.dtype != torch.int32:
        idx = idx.to(torch.int32)
    idx = idx.contiguous()
    if x.dtype not in _DT:
        raise TypeError(f"unsupported dtype {x.dtype}")
    B, N, C = x.shape
    if idx.shape[0] != B * groups or idx.shape[1] != N:
        raise ValueError(f"idx shape {tuple(idx.shape)} does not match B*G={B * groups}, N={N}")
    return _MRAggregate.apply(x, y, idx, groups)


# ----------------------------------------------------------------------------------------
# key pooling (avg_pool2d(x, r, r) on the token-major layout)
# ----------------------------------------------------------------------------------------
class _PoolKeys(torch.autograd.Function):
    @staticmethod
    @_guard
    def forward(ctx, x, H, W, r):
        lib = _lib.load()
        x = _token_major(x)
        B, N, C = x.shape
        y = torch.empty((B, (H // r) * (W // r), C), dtype=x.dtype, device=x.device)
        rc = lib.gkg_pool_keys_fwd(x.data_ptr(), x.stride(0), x.stride(1), y.data_ptr(), B, H, W, C, r,
                                   _DT[x.dtype], _stream(x))
        _lib.check(rc, "gkg_pool_keys_fwd")
        ctx.meta = (B, H, W, C, r, x.dtype)
        return y

    @staticmethod
    @_guard
    def backward(ctx, grad_y):
        B, H, W, C, r, dtype = ctx.meta
        lib = _lib.load()
        grad_y = grad_y.to(dtype).contiguous()
        gx = torch.empty((B, H * W, C), dtype=dtype, device=grad_y.device)
        rc = lib.gkg_pool_keys_bwd(grad_y.data_ptr(), gx.data_ptr(), B, H, W, C, r, _DT[dtype], _stream(grad_y))
        _lib.check(rc, "gkg_pool_keys_bwd")
        return gx, None, None, None


def pool_keys(x, H, W, r):
    """Keys of the dynamic graph convolution: ``avg_pool2d(x, r, r)`` (torch_vertex.py:194-196) on a
    token-major ``x (B, H*W, C)``; returns ``(B, (H//r)*(W//r), C)``.  Differentiable w.r.t. x."""
    _require_cuda(x)
    if x.dtype not in _DT:
        raise TypeError(f"unsupported dtype {x.dtype}")
    if x.shape[1] != H * W:
        raise ValueError(f"x has {x.shape[1]} nodes, expected H*W = {H * W}")
    if H // r < 1 or W // r < 1:
        raise ValueError(f"pooling window {r} larger than the {H}x{W} map")
    return _PoolKeys.apply(x, int(H), int(W), int(r))


# ----------------------------------------------------------------------------------------
# grouped 1x1 FC (+ folded norm + activation) on the tensor cores, inference form
# ----------------------------------------------------------------------------------------
_ACT = {None: 0, "none": 0, "relu": 1, "gelu": 2}


def grouped_fc_supported(c2: int) -> bool:
    return bool(_lib.load().gkg_grouped_fc_supported(int(c2)))


def grouped_fc_weights(weight, scale=None, transpose=False):
    """Conv2d(2C, 2C, 1, groups=4) weight ``(2C, 2C/4, 1, 1)`` -> bf16 tensor-core operand order (K-major 8x8 core
    matrices, zero padded): ``[4][NP/8][KP/8][8][8]`` for narrow groups, ``[4][passes][NT/8][KP/8][8][8]`` for the
    wide ones (``NT = gkg_grouped_fc_pass_width``), KP = NP = ceil16(2C/4).  One kernel launch.
    ``scale`` (2C,), e.g. the folded batch-norm gain, multiplies the output channels in fp32 first;
    ``transpose`` packs the per-group transposed weights (the data gradient of the convolution)."""
    c2, cg = weight.shape[0], weight.shape[1]
    if c2 != 4 * cg:
        raise ValueError(f"expected a groups=4 1x1 conv weight, got {tuple(weight.shape)}")
    _require_cuda(weight, scale)
    lib = _lib.load()
    nt = lib.gkg_grouped_fc_pass_width(int(c2))
    if nt < 0:
        raise ValueError(f"no grouped-FC kernel for 2C = {c2}")
    kp = (cg + 15) // 16 * 16
    rows = kp if nt == 0 else (kp + nt - 1) // nt * nt
    w = weight.detach().reshape(c2, cg).float().contiguous()
    sc = None if scale is None else scale.detach().reshape(c2).float().contiguous()
    out = torch.empty(4 * rows * kp, dtype=torch.bfloat16, device=weight.device)
    with torch.cuda.device(weight.device):
        rc = lib.gkg_grouped_fc_pack_weights(w.data_ptr(), None if sc is None else sc.data_ptr(), out.data_ptr(), c2,
                                             int(bool(transpose)), _stream(weight))
    _lib.check(rc, "gkg_grouped_fc_pack_weights")
    return out


@_guard
def grouped_fc(x, w_op, shift, act="gelu"):
    """``act(conv1x1_groups4(x; scale * W) + shift)`` on token-major bf16 ``x (..., 2C)``: the eval-mode
    ``BasicConv([2C, 2C])`` of the reference (torch_nn.py:57-81) in one pass.  ``w_op`` from
    :func:`grouped_fc_weights` (batch-norm gain folded in); ``shift`` fp32 ``(2C,)`` carries the conv
    bias and the rest of the norm."""
    _require_cuda(x, w_op, shift)
    if x.dtype != torch.bfloat16:
        raise TypeError("grouped_fc runs on bf16 activations")
    if act not in _ACT:
        raise NotImplementedError(f"activation [{act}] is not fused")
    c2 = x.shape[-1]
    x = x.contiguous()
    out = torch.empty_like(x)
    rows = x.numel() // c2
    rc = _lib.load().gkg_grouped_fc_fwd(x.data_ptr(), w_op.data_ptr(), shift.data_ptr(), out.data_ptr(), rows, c2,
                                        _ACT[act], _stream(x))
    _lib.check(rc, "gkg_grouped_fc_fwd")
    return out


class _GroupedFC(torch.autograd.Function):
    """Training form of the grouped 1x1 FC: ``x (.., 2C) bf16 -> conv1x1_groups4(x) + bias`` (pre-norm).
    Forward, the data gradient (the same product with the transposed weights) and the weight gradient
    (a (CG x CG) reduction over all rows per group) run on tcgen05 kernels."""

    @staticmethod
    @_guard
    def forward(ctx, x, weight, bias):
        c2 = x.shape[-1]
        shift = bias.detach().float() if bias is not None else torch.zeros(c2, device=x.device)
        out = grouped_fc(x, grouped_fc_weights(weight.detach()), shift.contiguous(), None)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    @_guard
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        c2 = x.shape[-1]
        cg = c2 // 4
        go = grad_out.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            zeros = torch.zeros(c2, device=x.device, dtype=torch.float32)
            gx = grouped_fc(go.to(torch.bfloat16), grouped_fc_weights(weight, transpose=True), zeros, None)
        if ctx.needs_input_grad[1]:
            # per conv group (CG x R) . (R x CG) on the tensor cores, split over the rows, fp32 accumulation
            gw32 = torch.zeros((4, cg, cg), dtype=torch.float32, device=x.device)
            gob = go.to(torch.bfloat16).reshape(-1, c2)
            xb = x.reshape(-1, c2)
            rc = _lib.load().gkg_grouped_fc_wgrad(gob.data_ptr(), xb.data_ptr(), gw32.data_ptr(), gob.shape[0], c2,
                                                  _stream(x))
            _lib.check(rc, "gkg_grouped_fc_wgrad")
            gw = gw32.reshape(c2, cg, 1, 1).to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            go2 = go.reshape(-1, c2)
            if go2.dtype in _DT and go2.data_ptr() % 16 == 0:
                gb = column_sum(go2)                  # one pass, fp32 accumulation (was: an fp32 copy + ATen's reduction)
            else:
                gb = go2.float().sum(0)
        return gx, gw, gb


def grouped_fc_train(x, weight, bias):
    """Differentiable ``conv1x1_groups4(x) + bias`` on token-major bf16 ``x (..., 2C)`` (pre-norm output)."""
    _require_cuda(x, weight)
    return _GroupedFC.apply(x, weight, bias)


# ----------------------------------------------------------------------------------------
# label-query head: per-label scores and the two multi-label losses
# ----------------------------------------------------------------------------------------
class _LabelScore(torch.autograd.Function):
    @staticmethod
    def forward(ctx, L, gap, W1, b1, W2, b2):
        lib = _lib.load()
        Lf, gf = L.float().contiguous(), gap.float().contiguous()
        w1, w2 = W1.float().contiguous(), W2.float().contiguous()
        B, n, C = Lf.shape
        score = torch.empty((B, n), dtype=torch.float32, device=L.device)
        with torch.cuda.device(L.device):
            rc = lib.gkg_label_score_fwd(Lf.data_ptr(), gf.data_ptr(), w1.data_ptr(), b1.float().contiguous().data_ptr(),
                                         w2.data_ptr(), b2.float().contiguous().data_ptr(), score.data_ptr(), B, n, C,
                                         _stream(L))
        _lib.check(rc, "gkg_label_score_fwd")
        ctx.save_for_backward(Lf, gf, w1, w2)
        ctx.dtypes = (L.dtype, gap.dtype, W1.dtype, b1.dtype, W2.dtype, b2.dtype)
        return score

    @staticmethod
    def backward(ctx, ds):
        Lf, gf, w1, w2 = ctx.saved_tensors
        B, n, C = Lf.shape
        dev = Lf.device
        ds = ds.float().contiguous()
        dL = torch.empty_like(Lf)
        dg = torch.empty_like(gf)
        dW1, dW2 = torch.empty_like(w1), torch.empty_like(w2)
        db1 = torch.empty(n, dtype=torch.float32, device=dev)
        db2 = torch.empty(n, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _lib.load().gkg_label_score_bwd(ds.data_ptr(), Lf.data_ptr(), gf.data_ptr(), w1.data_ptr(), w2.data_ptr(),
                                                 dL.data_ptr(), dg.data_ptr(), dW1.data_ptr(), dW2.data_ptr(), db1.data_ptr(),
                                                 db2.data_ptr(), B, n, C, _stream(ds))
        _lib.check(rc, "gkg_label_score_bwd")
        dt = ctx.dtypes
        return dL.to(dt[0]), dg.to(dt[1]), dW1.to(dt[2]), db1.to(dt[3]), dW2.to(dt[4]), db2.to(dt[5])


def label_score(label_emb, gap, fc1_weight, fc1_bias, fc2_weight, fc2_bias):
    """``score[b, i] = fc1.weight[i] . L[b, i] + fc1.bias[i] + fc2(gap)[b, i]`` (label_query_head.py:49-57)
    as one row-dot kernel, fp32, differentiable."""
    _require_cuda(label_emb, gap, fc1_weight, fc2_weight)
    return _LabelScore.apply(label_emb, gap, fc1_weight, fc1_bias, fc2_weight, fc2_bias)


class _MultiLabelLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, score, target, gamma_pos, gamma_neg, clip, eps, smooth):
        s = score.float().contiguous()
        t = target.to(torch.float32).contiguous()
        sums = torch.empty(2, dtype=torch.float32, device=s.device)
        d_asl, d_bce = torch.empty_like(s), torch.empty_like(s)
        with torch.cuda.device(s.device):
            rc = _lib.load().gkg_multilabel_loss(s.data_ptr(), t.data_ptr(), sums.data_ptr(), d_asl.data_ptr(),
                                                 d_bce.data_ptr(), s.numel(), float(gamma_pos), float(gamma_neg),
                                                 float(clip or 0.0), float(eps), float(smooth), _stream(s))
        _lib.check(rc, "gkg_multilabel_loss")
        ctx.save_for_backward(d_asl, d_bce)
        ctx.dtype = score.dtype
        return sums[0], sums[1]

    @staticmethod
    def backward(ctx, g_asl, g_bce):
        d_asl, d_bce = ctx.saved_tensors
        return (d_asl * g_asl + d_bce * g_bce).to(ctx.dtype), None, None, None, None, None, None


def multilabel_losses(score, target, gamma_pos=0.0, gamma_neg=2.0, clip=0.05, eps=1e-8, smooth=0.1):
    """(sum of AsymmetricLoss terms, sum of label-smoothed BCE-with-logits terms) over all entries of ``score`` --
    the two losses of LabelQueryHead.forward_train before the division by the batch -- in one kernel."""
    _require_cuda(score, target)
    return _MultiLabelLoss.apply(score, target, gamma_pos, gamma_neg, clip, eps, smooth)


# ----------------------------------------------------------------------------------------
# neighbour gather / neighbour sum (the graph-convolution variants other than max-relative)
# ----------------------------------------------------------------------------------------
class _NeighborGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, idx, groups, reduce_sum):
        lib = _lib.load()
        y = _token_major(y)
        B, M, C = y.shape
        P, N, k = idx.shape
        D = C // groups
        shape = (B, N, C) if reduce_sum else (B, N, k, C)
        out = torch.empty(shape, dtype=y.dtype, device=y.device)
        fn = lib.gkg_neighbor_sum_fwd if reduce_sum else lib.gkg_neighbor_gather_fwd
        with torch.cuda.device(y.device):
            rc = fn(y.data_ptr(), y.stride(0), y.stride(1), idx.data_ptr(), out.data_ptr(), B, groups, N, M, D, k,
                    _DT[y.dtype], _stream(y))
        _lib.check(rc, "gkg_neighbor_sum_fwd" if reduce_sum else "gkg_neighbor_gather_fwd")
        ctx.save_for_backward(idx)
        ctx.meta = (B, groups, N, M, D, k, y.dtype, reduce_sum)
        return out

    @staticmethod
    def backward(ctx, grad):
        (idx,) = ctx.saved_tensors
        B, G, N, M, D, k, dtype, reduce_sum = ctx.meta
        grad = grad.contiguous().to(dtype)
        gy = torch.zeros((B, M, G * D), dtype=torch.float32, device=grad.device)
        lib = _lib.load()
        fn = lib.gkg_neighbor_sum_bwd if reduce_sum else lib.gkg_neighbor_gather_bwd
        with torch.cuda.device(grad.device):
            rc = fn(grad.data_ptr(), idx.data_ptr(), gy.data_ptr(), B, G, N, M, D, k, _DT[dtype], _stream(grad))
        _lib.check(rc, "gkg_neighbor_gather_bwd")
        return gy.to(dtype), None, None, None


def _check_gather(y, idx, groups):
    _require_cuda(y, idx)
    if y.dtype not in _DT:
        raise TypeError(f"unsupported dtype {y.dtype}")
    if idx.dtype != torch.int32:
        idx = idx.to(torch.int32)
    idx = idx.contiguous()
    if y.shape[2] % groups or idx.shape[0] != y.shape[0] * groups:
        raise ValueError(f"idx {tuple(idx.shape)} does not match B*G = {y.shape[0] * groups}")
    return idx


def gather_neighbors(y, idx, *, groups=1):
    """``out[b, n, j, c] = y[b, idx[b*G + c//D, n, j], c]`` -- batched_index_select (torch_nn.py:84-105) on the
    token-major layout: y (B, M, C), idx int32 (B*G, N, k) -> (B, N, k, C).  Differentiable w.r.t. y."""
    return _NeighborGather.apply(y, _check_gather(y, idx, groups), groups, False)


def sum_neighbors(y, idx, *, groups=1):
    """``out[b, n, c] = sum_j y[b, idx[b*G + c//D, n, j], c]`` (GINConv2d, torch_vertex.py:143-149) without the
    gathered (B, N, k, C) tensor.  Differentiable w.r.t. y."""
    return _NeighborGather.apply(y, _check_gather(y, idx, groups), groups, True)


# ---------------------------------------------------------------------------------------------------------
# training-mode batch norm: native statistics / backward reductions (csrc/batch_norm.cu), ATen elementwise halves
# ---------------------------------------------------------------------------------------------------------
_COUNT_CACHE = {}


def _rows_of(x):
    """(rows, C) contiguous view of a channels-last (B, C, H, W) or token-major (..., C) activation, or None."""
    if x.dim() == 4:
        t = x.permute(0, 2, 3, 1)
        return t.reshape(-1, x.shape[1]) if t.is_contiguous() else None
    return None


def batch_norm_native_ok(x):
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 4 and x.numel() > 0):
        return False
    if x.dtype not in (torch.bfloat16, torch.float32):
        return False
    if x.shape[1] % (8 if x.dtype == torch.bfloat16 else 4) != 0 or x.shape[0] * x.shape[2] * x.shape[3] < 2:
        return False
    return x.permute(0, 2, 3, 1).is_contiguous() and x.data_ptr() % 16 == 0


def _sync_world(group):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


class _BatchNormTrain(torch.autograd.Function):
    """y = act(batch_norm(x)) with batch statistics (torch_nn.py:32-42 -> nn.BatchNorm2d / nn.SyncBatchNorm training
    forward, optionally followed by the stack's nn.GELU): statistics and backward reductions by gkg_bn_stats /
    gkg_bn_backward_reduce.
    act None: elementwise passes by ATen's batch_norm_elemt / batch_norm_backward_elemt (the decomposition
    SyncBatchNorm uses).  act "gelu": gkg_bn_act_forward / gkg_bn_act_backward -- the activation rides on the
    normalisation pass, its derivative is saved in the feature dtype and the backward passes stream it.
    sync: statistics over all data-parallel ranks (the reference's SyncBN, torch_nn.py:8): one all-gather of the
    per-rank (mean, var) in the forward (equal row counts per rank: Chan's merge with equal weights), one all-reduce
    of (sum g, sum g (x - mean)) in the backward -- two small collectives per norm, both capturable."""

    @staticmethod
    @_guard
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, act, group_sync):
        lib = _lib.load()
        x2 = _rows_of(x)
        rows, C = x2.shape
        sync, group = group_sync
        world = _sync_world(group) if sync else 1
        mean = torch.empty(C, dtype=torch.float32, device=x.device)
        invstd = torch.empty(C, dtype=torch.float32, device=x.device)
        ws = _workspace(x.device, lib.gkg_bn_workspace_bytes(rows, C))
        local_running = world == 1
        rc = lib.gkg_bn_stats(x2.data_ptr(), rows, C, _DT[x.dtype], float(eps), float(momentum), mean.data_ptr(),
                              invstd.data_ptr(),
                              None if (running_mean is None or not local_running) else running_mean.data_ptr(),
                              None if (running_var is None or not local_running) else running_var.data_ptr(),
                              ws.data_ptr(), ws.numel(), _stream(x))
        _lib.check(rc, "gkg_bn_stats")
        if world > 1:
            import torch.distributed as dist
            local = torch.stack((mean, invstd.pow(-2) - eps))                  # (2, C): mean, biased variance of this rank
            allst = torch.empty((world, 2, C), dtype=torch.float32, device=x.device)
            dist.all_gather_into_tensor(allst, local, group=group)
            mean = allst[:, 0].mean(0)
            var = allst[:, 1].mean(0) + (allst[:, 0] - mean).square().mean(0)
            invstd = torch.rsqrt(var + eps)
            if running_mean is not None:
                n = rows * world
                running_mean.mul_(1.0 - momentum).add_(mean, alpha=momentum)
                running_var.mul_(1.0 - momentum).add_(var, alpha=momentum * n / max(n - 1, 1))
        ctx.act, ctx.world, ctx.group = act, world, group
        if act is None:
            y = torch.batch_norm_elemt(x, weight, bias, mean, invstd, eps)
            ctx.save_for_backward(x, weight, mean, invstd)
            return y
        y = torch.empty_like(x)                                   # same (channels-last) strides
        # gelu'(z) saved in the feature dtype (what the unfused path keeps is the norm output, same size): the
        # backward passes then stream dy, x and this tensor without evaluating the derivative again
        dact = torch.empty_like(x) if x.requires_grad or weight.requires_grad else None
        rc = lib.gkg_bn_act_forward(x2.data_ptr(), mean.data_ptr(), invstd.data_ptr(), weight.data_ptr(),
                                    bias.data_ptr(), rows, C, _DT[x.dtype], 2, y.data_ptr(),
                                    None if dact is None else dact.data_ptr(), _stream(x))
        _lib.check(rc, "gkg_bn_act_forward")
        ctx.save_for_backward(x, weight, mean, invstd, bias, dact)
        return y

    @staticmethod
    @_guard
    def backward(ctx, dy):
        lib = _lib.load()
        x, weight, mean, invstd = ctx.saved_tensors[:4]
        if not dy.permute(0, 2, 3, 1).is_contiguous():
            dy = dy.contiguous(memory_format=torch.channels_last)
        dy = dy.to(x.dtype)
        x2, g2 = _rows_of(x), _rows_of(dy)
        rows, C = x2.shape
        world = ctx.world
        out = torch.empty(4, C, dtype=torch.float32, device=x.device)       # sum_dy, sum_dy_xmu, grad_weight, grad_bias
        ws = _workspace(x.device, lib.gkg_bn_workspace_bytes(rows, C))
        if ctx.act is not None:
            bias, dact = ctx.saved_tensors[4], ctx.saved_tensors[5]
            dx = torch.empty_like(x)
            dptr = None if dact is None else dact.data_ptr()
            if world == 1:
                rc = lib.gkg_bn_act_backward(g2.data_ptr(), x2.data_ptr(), dptr, mean.data_ptr(), invstd.data_ptr(),
                                             weight.data_ptr(), bias.data_ptr(), rows, C, _DT[x.dtype], 2, dx.data_ptr(),
                                             out[2].data_ptr(), out[3].data_ptr(), ws.data_ptr(), ws.numel(), _stream(x))
                _lib.check(rc, "gkg_bn_act_backward")
            else:
                import torch.distributed as dist
                rc = lib.gkg_bn_act_backward_reduce(g2.data_ptr(), x2.data_ptr(), dptr, mean.data_ptr(), invstd.data_ptr(),
                                                    weight.data_ptr(), bias.data_ptr(), rows, C, _DT[x.dtype], 2,
                                                    out[0].data_ptr(), out[2].data_ptr(), out[3].data_ptr(),
                                                    ws.data_ptr(), ws.numel(), _stream(x))
                _lib.check(rc, "gkg_bn_act_backward_reduce")
                dist.all_reduce(out[:2], group=ctx.group)                     # (sum g, sum g (x - mean)) over all ranks
                rc = lib.gkg_bn_act_backward_elemt(g2.data_ptr(), x2.data_ptr(), dptr, mean.data_ptr(), invstd.data_ptr(),
                                                   weight.data_ptr(), bias.data_ptr(), out[0].data_ptr(), rows,
                                                   rows * world, C, _DT[x.dtype], 2, dx.data_ptr(), _stream(x))
                _lib.check(rc, "gkg_bn_act_backward_elemt")
        else:
            rc = lib.gkg_bn_backward_reduce(g2.data_ptr(), x2.data_ptr(), mean.data_ptr(), invstd.data_ptr(), rows, C,
                                            _DT[x.dtype], out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                            out[3].data_ptr(), ws.data_ptr(), ws.numel(), _stream(x))
            _lib.check(rc, "gkg_bn_backward_reduce")
            if world > 1:
                import torch.distributed as dist
                dist.all_reduce(out[:2], group=ctx.group)
            key = (x.device.index, rows, world)
            count = _COUNT_CACHE.get(key)
            if count is None:
                count = _COUNT_CACHE[key] = torch.full((world,), rows, dtype=torch.int32, device=x.device)
            dx = None
            if ctx.needs_input_grad[0]:
                dx = torch.batch_norm_backward_elemt(dy, x, mean, invstd, weight, out[0], out[1], count)
        gw = out[2].to(weight.dtype) if ctx.needs_input_grad[1] else None
        gb = out[3].to(weight.dtype) if ctx.needs_input_grad[2] else None
        return dx, gw, gb, None, None, None, None, None, None


def batch_norm_train(x, weight, bias, running_mean, running_var, momentum, eps, act=None, sync=False, group=None):
    """Training-mode batch norm of a channels-last (B, C, H, W) CUDA activation, optionally fused with the GELU that
    follows it (act="gelu"); updates the running statistics in place like nn.BatchNorm2d.  sync=True: statistics over
    all ranks of ``group`` (nn.SyncBatchNorm semantics; every rank must hold the same number of rows).  Raises on
    anything the kernels do not take (callers check batch_norm_native_ok)."""
    _require_cuda(x)
    if not batch_norm_native_ok(x):
        raise ValueError("batch_norm_train: needs a channels-last (B, C, H, W) bf16 / fp32 CUDA tensor, C % 8 (4) == 0")
    if act not in (None, "gelu"):
        raise ValueError(f"batch_norm_train: activation {act!r} (None or 'gelu')")
    if act is not None and (weight is None or bias is None or weight.dtype != torch.float32):
        raise ValueError("batch_norm_train: the fused activation needs fp32 affine parameters")
    return _BatchNormTrain.apply(x, weight, bias, running_mean, running_var, momentum, eps, act, (bool(sync), group))


# ---------------------------------------------------------------------------------------------------------
# token-major 1x1 convolution (fc1 / fc2 / FFN of the graph blocks) with the bias gradient by gkg_column_sum
# ---------------------------------------------------------------------------------------------------------
@_guard
def column_sum(x2):
    """fp32 column sums of a contiguous (rows, C) CUDA tensor (bf16 / fp32, C % 8 (4) == 0)."""
    _require_cuda(x2)
    lib = _lib.load()
    rows, C = x2.shape
    out = torch.empty(C, dtype=torch.float32, device=x2.device)
    ws = _workspace(x2.device, lib.gkg_bn_workspace_bytes(rows, C))
    rc = lib.gkg_column_sum(x2.data_ptr(), rows, C, _DT[x2.dtype], out.data_ptr(), ws.data_ptr(), ws.numel(), _stream(x2))
    _lib.check(rc, "gkg_column_sum")
    return out


class _Conv1x1(torch.autograd.Function):
    """y = conv2d(x, w, b) for a 1x1, stride-1 convolution on a channels-last activation = one token-major GEMM
    (bias in the epilogue).  Backward: the two GEMMs autograd would run, and the bias gradient as one pass of
    gkg_column_sum instead of ATen's generic reduction (3.8 ms of a GKGNet-576 training step)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        xt = x.permute(0, 2, 3, 1)                               # (B, H, W, Cin), contiguous for channels-last x
        w = weight.view(weight.shape[0], -1).to(xt.dtype)
        y = torch.nn.functional.linear(xt, w, None if bias is None else bias.to(xt.dtype))
        ctx.save_for_backward(xt, w)
        ctx.has_bias = bias is not None
        ctx.wshape, ctx.wdtype = weight.shape, weight.dtype
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        xt, w = ctx.saved_tensors
        g = dy.permute(0, 2, 3, 1)
        if not g.is_contiguous():
            g = g.contiguous()
        g2 = g.reshape(-1, g.shape[-1]).to(xt.dtype)
        x2 = xt.reshape(-1, xt.shape[-1])
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = (g2 @ w).view(xt.shape).permute(0, 3, 1, 2)
        if ctx.needs_input_grad[1]:
            dw = (g2.t() @ x2).to(ctx.wdtype).view(ctx.wshape)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            C = g2.shape[1]
            if g2.is_cuda and C % (8 if g2.dtype == torch.bfloat16 else 4) == 0 and g2.dtype in _DT and g2.data_ptr() % 16 == 0:
                db = column_sum(g2).to(ctx.wdtype)
            else:
                db = g2.float().sum(0).to(ctx.wdtype)
        return dx, dw, db


def conv1x1(x, weight, bias):
    """Channels-last 1x1 convolution as a token-major GEMM (autocast-aware: runs in the activation's dtype)."""
    return _Conv1x1.apply(x, weight, bias)

