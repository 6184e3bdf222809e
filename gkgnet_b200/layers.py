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


class BatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d (same parameters, buffers and state-dict keys) whose training forward on channels-last CUDA
    activations takes its statistics and backward reductions from the native kernels (ops.batch_norm_train:
    SURVEY 8(f) rank 1); everything else -- eval mode, CPU, NCHW-contiguous input, odd widths -- is the stock module."""

    native = True      # class-wide switch (tests compare against the stock path)

    def takes_native(self, x):
        if not (self.native and self.training and self.affine and self.track_running_stats and self.momentum is not None
                and isinstance(x, torch.Tensor) and x.is_cuda):
            return False
        from . import ops
        return ops.batch_norm_native_ok(x)

    def forward(self, x, act=None):
        """act="gelu": the exact GELU that follows this norm in the stack, applied in the same pass (native path only;
        callers check takes_native first)."""
        if self.takes_native(x):
            from . import ops
            self.num_batches_tracked.add_(1)
            return ops.batch_norm_train(x, self.weight, self.bias, self.running_mean, self.running_var,
                                        self.momentum, self.eps, act)
        if act is not None:
            raise ValueError("fused activation needs the native path")
        return super().forward(x)


class SyncBatchNorm(nn.SyncBatchNorm):
    """nn.SyncBatchNorm (the reference's norm: torch_nn.py:8, gkgnet.py:23; same parameters, buffers, state-dict keys)
    whose training forward on channels-last CUDA activations runs the native kernels with the statistics exchanged by
    one all-gather (forward) and one all-reduce (backward) of per-channel vectors.  Every rank must hold the same number
    of rows (the benchmark's fixed per-GPU batch; the stock module also handles ragged batches and is used otherwise)."""

    native = True

    def takes_native(self, x):
        if not (self.native and self.training and self.affine and self.track_running_stats and self.momentum is not None
                and isinstance(x, torch.Tensor) and x.is_cuda):
            return False
        from . import ops
        return ops.batch_norm_native_ok(x)

    def forward(self, x, act=None):
        if self.takes_native(x):
            from . import ops
            self.num_batches_tracked.add_(1)
            return ops.batch_norm_train(x, self.weight, self.bias, self.running_mean, self.running_var,
                                        self.momentum, self.eps, act, sync=True, group=self.process_group)
        if act is not None:
            raise ValueError("fused activation needs the native path")
        return super().forward(x)


def run_modules(mods, x):
    """Run a conv -> norm -> act stack as written (training / gradient passes), with two substitutions that keep the
    arithmetic: a 1x1 convolution on channels-last CUDA activations is the token-major GEMM it is (ops.conv1x1), and a
    native batch norm followed by the exact nn.GELU runs as one pass (ops.batch_norm_train(act="gelu"))."""
    mods = list(mods)
    i = 0
    while i < len(mods):
        m = mods[i]
        nxt = mods[i + 1] if i + 1 < len(mods) else None
        if (isinstance(m, nn.Conv2d) and m.kernel_size == (1, 1) and m.stride == (1, 1) and m.padding == (0, 0)
                and m.groups == 1 and x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and x.is_cuda):
            if torch.is_autocast_enabled():
                x = x.to(torch.get_autocast_dtype("cuda"))
            from . import ops
            x = ops.conv1x1(x, m.weight, m.bias)
        elif (isinstance(m, (BatchNorm2d, SyncBatchNorm)) and isinstance(nxt, nn.GELU) and getattr(nxt, "approximate", "none") == "none"
              and m.weight is not None and m.weight.dtype == torch.float32 and m.takes_native(x)):
            x = m(x, act="gelu")
            i += 1
        else:
            x = m(x)
        i += 1
    return x


def build_norm_layer(cfg, num_features, postfix=""):
    """Stand-in for mmcv.cnn.build_norm_layer with the two types the reference uses."""
    kind = cfg.get("type", "SyncBN")
    if kind == "SyncBN":
        layer = SyncBatchNorm(num_features)
    elif kind == "BN":
        layer = BatchNorm2d(num_features)
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


class FoldedSequential(nn.Sequential):
    """``nn.Sequential`` (same sub-module names, hence the same state-dict keys as the reference's conv -> norm
    stacks: torch_vertex.py:290-306, gkgnet.py:52-65,82-95,108-112) that, in eval mode without autograd, folds
    every ``Conv2d -> BatchNorm`` pair into one convolution:  w' = w * g / sqrt(var + eps),
    b' = (b - mean) * g / sqrt(var + eps) + beta.  Training and gradient passes run the modules as written."""

    def forward(self, x):
        if self.training or torch.is_grad_enabled():
            return run_modules(self, x)
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            if (isinstance(m, nn.Conv2d) and isinstance(nxt, (nn.BatchNorm2d, nn.SyncBatchNorm))
                    and nxt.track_running_stats and nxt.running_mean is not None and not nxt.training):
                w, b = self._folded(i, m, nxt, x.dtype if not torch.is_autocast_enabled() else torch.get_autocast_dtype("cuda"))
                x = x.to(w.dtype)
                if (m.kernel_size == (1, 1) and m.stride == (1, 1) and m.padding == (0, 0) and m.groups == 1
                        and x.is_contiguous(memory_format=torch.channels_last)):
                    # token-major 1x1 conv == one GEMM with the bias in its epilogue (conv2d adds it in a second pass)
                    x = torch.nn.functional.linear(x.permute(0, 2, 3, 1), w.view(w.shape[0], -1), b).permute(0, 3, 1, 2)
                else:
                    x = torch.nn.functional.conv2d(x, w, b, m.stride, m.padding, m.dilation, m.groups)
                i += 2
            else:
                x = m(x)
                i += 1
        return x

    def _folded(self, i, conv, bn, dtype):
        tensors = [conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        key = (dtype,) + tuple((t.data_ptr(), t._version) for t in tensors if t is not None)
        cache = self.__dict__.setdefault("_fold_cache", {})
        hit = cache.get(i)
        if hit is None or hit[0] != key:
            with torch.no_grad():
                scale = (bn.weight.float() if bn.weight is not None else 1.0) / torch.sqrt(bn.running_var.float() + bn.eps)
                w = (conv.weight.float() * scale.view(-1, 1, 1, 1)).to(dtype)
                b0 = conv.bias.float() if conv.bias is not None else torch.zeros_like(bn.running_mean, dtype=torch.float32)
                b = (b0 - bn.running_mean.float()) * scale + (bn.bias.float() if bn.bias is not None else 0.0)
                hit = (key, w.contiguous(), b.to(dtype).contiguous())
            cache[i] = hit
        return hit[1], hit[2]


class BasicConv(FoldedSequential):
    """Stack of grouped(4) 1x1 convs, each followed by norm / act / dropout
    (torch_nn.py:57-81).  Sub-module order -- and therefore the state-dict keys
    ``0.weight, 0.bias, 1.<bn>`` -- matches the reference.  Widths the tensor-core grouped FC does not take
    (stages 3-4) run here; in eval mode the batch norm is folded into the grouped convolution."""

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


class _ScaledResidual(torch.autograd.Function):
    """shortcut + x * scale (scale: per-sample, broadcast) in one pass; the gradient of the low-precision branch is
    written in its own dtype by the same kernel that scales it (autograd's addcmul backward multiplies in the residual
    stream's fp32 and casts in a second pass)."""

    @staticmethod
    def forward(ctx, x, shortcut, scale):
        ctx.save_for_backward(scale)
        ctx.xdtype = x.dtype
        return torch.addcmul(shortcut, x, scale)

    @staticmethod
    def backward(ctx, g):
        (scale,) = ctx.saved_tensors
        gx = None
        if ctx.needs_input_grad[0]:
            gx = torch.empty_like(g, dtype=ctx.xdtype)
            torch.mul(g, scale, out=gx)
        return gx, (g if ctx.needs_input_grad[1] else None), None


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

    def add_residual(self, x, shortcut):
        """``self(x) + shortcut`` in one elementwise pass (the per-sample keep / scale mask rides on the residual add)."""
        if self.drop_prob == 0.0 or not self.training:
            return x + shortcut
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.dim() - 1)
        mask = torch.empty(shape, dtype=torch.promote_types(x.dtype, shortcut.dtype), device=x.device)
        return _ScaledResidual.apply(x, shortcut, mask.bernoulli_(keep).div_(keep))

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
