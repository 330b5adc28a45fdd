"""Minimal stand-ins for the mmcv pieces the GeoMAE configs and model code rely on
(mmcv is not a dependency here): ``Registry`` with ``register_module``/``build`` and a
``Config`` that loads the reference's plain-Python config files including ``_base_``
inheritance (mmcv.Config.fromfile semantics: dict-merge, child wins; reference
tools/train.py:101-103).  Registry names follow mmdet3d/models/builder.py:5-14."""
from __future__ import annotations

import os


class Registry:
    def __init__(self, name):
        self.name = name
        self._modules = {}

    def register_module(self, name=None, force=False, module=None):
        if isinstance(name, type):      # used as a bare decorator
            self._modules[name.__name__] = name
            return name

        def deco(cls):
            self._modules[name or cls.__name__] = cls
            return cls
        return deco(module) if module is not None else deco

    def get(self, key):
        return self._modules.get(key)

    def build(self, cfg, default_args=None):
        if not isinstance(cfg, dict) or "type" not in cfg:
            raise TypeError(f"{self.name}: cfg must be a dict with a 'type' key, got {cfg!r}")
        args = dict(cfg)
        for k, v in (default_args or {}).items():
            args.setdefault(k, v)
        kind = args.pop("type")
        cls = self._modules.get(kind) if isinstance(kind, str) else kind
        if cls is None:
            raise KeyError(f"{kind} is not in the {self.name} registry")
        return cls(**args)


DETECTORS = Registry("detector")
BACKBONES = Registry("backbone")
VOXEL_ENCODERS = Registry("voxel_encoder")
MIDDLE_ENCODERS = VOXEL_ENCODERS
LOSSES = Registry("loss")
NORM_LAYERS = Registry("norm layer")


def build_detector(cfg, train_cfg=None, test_cfg=None):
    return DETECTORS.build(cfg, default_args=dict(train_cfg=train_cfg, test_cfg=test_cfg))


build_model = build_detector


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_voxel_encoder(cfg):
    return VOXEL_ENCODERS.build(cfg)


def build_loss(cfg):
    return LOSSES.build(cfg)


def build_norm_layer(cfg, num_features, postfix=""):
    """mmcv.cnn.build_norm_layer: returns (name, layer)."""
    args = dict(cfg)
    kind = args.pop("type")
    args.pop("requires_grad", None)
    cls = NORM_LAYERS.get(kind)
    if cls is None:
        raise KeyError(f"{kind} is not a registered norm layer")
    return "bn" + str(postfix), cls(num_features, **args)


def _merge(base: dict, child: dict) -> dict:
    out = dict(base)
    for k, v in child.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get("_delete_", False):
            out[k] = _merge(out[k], v)
        else:
            out[k] = {kk: vv for kk, vv in v.items() if kk != "_delete_"} if isinstance(v, dict) else v
    return out


class Config(dict):
    """``Config.fromfile(path)`` -> dict-like with attribute access for top-level keys."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    @staticmethod
    def _load(path):
        ns = {}
        with open(path) as f:
            exec(compile(f.read(), path, "exec"), ns)
        cfg = {k: v for k, v in ns.items() if not k.startswith("__") and not callable(v)
               and not isinstance(v, type(os))}
        bases = cfg.pop("_base_", [])
        if isinstance(bases, str):
            bases = [bases]
        merged = {}
        for b in bases:
            bpath = os.path.join(os.path.dirname(path), b)
            if os.path.exists(bpath):       # dataset/schedule bases are optional for the model path
                merged = _merge(merged, Config._load(bpath))
        return _merge(merged, cfg)

    @classmethod
    def fromfile(cls, path):
        return cls(cls._load(path))
