"""Plug-in mechanism.  The reference exposes its models through mmcv ``Registry`` objects
(mmcls/models/builder.py:6-14) and config dicts ``dict(type='GKGNet', ...)``; external
packages hook in with ``custom_imports`` (docs/en/tutorials/runtime.md:232-235).  When
mmcls/mmcv are importable we register our classes there (``force=True`` so they replace
the reference's classes of the same name); otherwise a minimal local registry with the
same ``register_module`` / ``build`` behaviour is used."""
from __future__ import annotations


class Registry:
    def __init__(self, name):
        self.name = name
        self._modules = {}

    def get(self, key):
        return self._modules.get(key)

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            key = name or cls.__name__
            if key in self._modules and not force:
                raise KeyError(f"{key} is already registered in {self.name}")
            self._modules[key] = cls
            return cls
        if module is not None:
            return _register(module)
        return _register

    def build(self, cfg, **default_args):
        if not isinstance(cfg, dict) or "type" not in cfg:
            raise TypeError(f'cfg must be a dict with the key "type", got {cfg!r}')
        args = dict(cfg)
        kind = args.pop("type")
        cls = self._modules.get(kind) if isinstance(kind, str) else kind
        if cls is None:
            raise KeyError(f"{kind} is not in the {self.name} registry")
        for k, v in default_args.items():
            args.setdefault(k, v)
        return cls(**args)


MODELS = Registry("models")
BACKBONES = NECKS = HEADS = LOSSES = CLASSIFIERS = MODELS


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)


def register_into_mmcls(*classes):
    """Best effort: also expose the classes through the real mmcls registry."""
    try:
        from mmcls.models.builder import MODELS as MMCLS_MODELS  # type: ignore
    except Exception:
        return False
    for cls in classes:
        MMCLS_MODELS.register_module(name=cls.__name__, force=True, module=cls)
    return True
