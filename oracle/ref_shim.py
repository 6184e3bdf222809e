"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference hot-path files.

The reference (jin-s13/GKGNet) is an mmcls fork that needs mmcv / timm / easydict,
none of which exist in this image.  Its graph hot path
(``mmcls/models/backbones/vig_model/*.py`` + ``mmcls/models/backbones/gkgnet.py``) is
plain PyTorch, so we pre-seed ``sys.modules`` with tiny stand-ins for the missing
imports and then import the reference files *from where they lie* (nothing is copied).

Used only by ``oracle/gen_golden.py`` (to create ``tests/golden/*.npz``) and by
CPU tests that cross-check ``oracle/gkg_oracle.py`` when the reference tree is present
(this container).  The GPU box has no ``/root/reference``: nothing under ``-m gpu``,
``smoke()`` or ``bench.py`` may call this module.

No reference code lives in this file; the stand-ins follow the *interfaces* the
reference imports at gkgnet.py:5-21, torch_vertex.py:3-12, torch_nn.py:3-7,
torch_edge.py:3-7, base_backbone.py:4.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np
import torch
from torch import nn

_DEFAULT_ROOTS = (os.environ.get("GKG_REF", ""), "/root/reference")


def find_reference_root():
    for root in _DEFAULT_ROOTS:
        if root and os.path.isfile(
                os.path.join(root, "mmcls/models/backbones/vig_model/torch_edge.py")):
            return root
    return None


class _StochasticDepth(nn.Module):
    """Stand-in for timm / mmcv ``DropPath`` (identity in eval or when p == 0)."""

    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x.div(keep) * mask


class _AttrDict(dict):
    """Stand-in for ``easydict.EasyDict``: attribute access, AttributeError on miss."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            self[k] = _AttrDict(v) if isinstance(v, dict) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e


def _build_norm_layer(cfg, num_features, postfix=""):
    kind = cfg["type"]
    if kind == "SyncBN":
        layer = nn.SyncBatchNorm(num_features)
    elif kind == "BN":
        layer = nn.BatchNorm2d(num_features)
    else:  # pragma: no cover - the reference only uses the two above
        raise KeyError(kind)
    for p in layer.parameters():
        p.requires_grad = cfg.get("requires_grad", True)
    return "bn" + str(postfix), layer


class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg

    def init_weights(self):
        pass


class _Registry:
    def register_module(self, *a, **k):
        def deco(cls):
            return cls
        return deco


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = None


def load_reference(root=None):
    """Return a namespace with the reference classes/functions of the hot path."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = root or find_reference_root()
    if root is None:
        raise FileNotFoundError("reference tree not found (set $GKG_REF)")
    if not hasattr(np, "float"):
        np.float = float  # pos_embed.py:74 uses the removed alias

    placeholder = lambda *a, **k: None  # noqa: E731
    _mod("mmcv", __version__="1.5.0")
    _mod("mmcv.cnn", ConvModule=placeholder, build_conv_layer=placeholder,
         build_norm_layer=_build_norm_layer, constant_init=placeholder)
    _mod("mmcv.cnn.bricks", DropPath=_StochasticDepth)
    _mod("mmcv.runner", BaseModule=_BaseModule)
    _mod("timm")
    _mod("timm.models")
    _mod("timm.models.layers", DropPath=_StochasticDepth)
    _mod("timm.data", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406),
         IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225))
    _mod("easydict", EasyDict=_AttrDict)

    base = os.path.join(root, "mmcls")
    for name, rel in (("mmcls", ""), ("mmcls.models", "models"),
                      ("mmcls.models.backbones", "models/backbones"),
                      ("mmcls.models.utils", "models/utils")):
        pkg = _mod(name)
        pkg.__path__ = [os.path.join(base, rel)]
    _mod("mmcls.models.builder", BACKBONES=_Registry(), HEADS=_Registry())

    vig = importlib.import_module("mmcls.models.backbones.vig_model")
    gkg = importlib.import_module("mmcls.models.backbones.gkgnet")
    ns = types.SimpleNamespace(root=root, vig=vig, gkgnet=gkg)
    for name in ("DenseDilatedKnnGraph", "DenseDilated", "MRConv2d", "BasicConv",
                 "batched_index_select", "DyGraphConv2d", "DyGraphConv2dMultiGroup",
                 "DyGraphLabel", "DyGraphLabelMultiGroup", "Grapher", "GrapherLabel",
                 "FFNLabel", "xy_dense_knn_matrix", "dense_knn_matrix",
                 "xy_pairwise_distance", "pairwise_distance", "part_pairwise_distance",
                 "get_2d_relative_pos_embed", "GraphConv2d", "EdgeConv2d", "GraphSAGE", "GINConv2d", "GraphAtten"):
        setattr(ns, name, getattr(vig, name))
    ns.GKGNet = gkg.GKGNet
    _loaded = ns
    return ns


class cpu_cuda_noop:
    """Context manager: make ``Tensor.cuda()`` the identity so the reference's
    hard-coded ``.cuda()`` (gkgnet.py:264) runs on a CPU-only box."""

    def __enter__(self):
        self._orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        return self

    def __exit__(self, *exc):
        torch.Tensor.cuda = self._orig
        return False
