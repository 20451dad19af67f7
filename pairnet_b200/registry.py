"""Minimal stand-in for the mmcv ``Registry`` / ``Config`` machinery the reference is built on
(mmcv-full 1.7.0 is not installable here).  Only what the hot path's drop-in surface needs:
type-string -> class lookup, ``build(cfg)``, python-dict configs with ``_base_`` inheritance and
``custom_imports`` (reference usage: ``tools/train.py:118-127,214-216``,
``configs/mask2former/pairnet.py:1,214-225``)."""
import copy
import importlib
import os


class ConfigDict(dict):
    """dict with attribute access (mmcv/addict style), recursively applied."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value

    def __deepcopy__(self, memo):
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def to_config(obj):
    if isinstance(obj, dict):
        return ConfigDict({k: to_config(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return type(obj)(to_config(v) for v in obj)
    return obj


class Registry:
    def __init__(self, name):
        self.name = name
        self._modules = {}

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            key = name or cls.__name__
            if key in self._modules and not force and self._modules[key] is not cls:
                raise KeyError(f"{key} is already registered in {self.name}")
            self._modules[key] = cls
            return cls

        if module is not None:
            return _register(module)
        return _register

    def get(self, key):
        return self._modules.get(key)

    def __contains__(self, key):
        return key in self._modules

    def build(self, cfg, default_args=None):
        if cfg is None:
            return None
        if not isinstance(cfg, dict) or "type" not in cfg:
            raise TypeError(f"{self.name}: cfg must be a dict with a 'type' key, got {cfg!r}")
        args = dict(cfg)
        typ = args.pop("type")
        cls = typ if isinstance(typ, type) else self.get(typ)
        if cls is None:
            raise KeyError(f"{typ} is not in the {self.name} registry")
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        return cls(**args)


DETECTORS = Registry("detector")
HEADS = Registry("head")
BACKBONES = Registry("backbone")
LOSSES = Registry("loss")
PLUGIN_LAYERS = Registry("plugin layer")
TRANSFORMER_LAYER_SEQUENCE = Registry("transformer layer sequence")
POSITIONAL_ENCODING = Registry("positional encoding")
BBOX_ASSIGNERS = Registry("bbox assigner")
BBOX_SAMPLERS = Registry("bbox sampler")
MATCH_COST = Registry("match cost")
DATASETS = Registry("dataset")
PIPELINES = Registry("pipeline")


def build_detector(cfg, train_cfg=None, test_cfg=None):
    """mmdet.models.build_detector (tools/train.py:214-216)."""
    return DETECTORS.build(cfg, default_args=dict(train_cfg=train_cfg, test_cfg=test_cfg))


def build_head(cfg):
    return HEADS.build(cfg)


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_loss(cfg):
    return LOSSES.build(cfg)


def build_transformer_layer_sequence(cfg):
    return TRANSFORMER_LAYER_SEQUENCE.build(cfg)


def build_positional_encoding(cfg):
    return POSITIONAL_ENCODING.build(cfg)


def build_plugin_layer(cfg):
    layer = PLUGIN_LAYERS.build(cfg)
    return type(layer).__name__, layer


def build_assigner(cfg):
    return BBOX_ASSIGNERS.build(cfg)


def build_sampler(cfg, **default_args):
    return BBOX_SAMPLERS.build(cfg, default_args=default_args)


# ------------------------------------------------------------------------------------------- Config
def _merge(base, new):
    out = dict(base)
    for k, v in new.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict) and not v.get("_delete_", False):
            out[k] = _merge(out[k], v)
        else:
            if isinstance(v, dict):
                v = {kk: vv for kk, vv in v.items() if kk != "_delete_"}
            out[k] = v
    return out


def _load_py(path):
    scope = {"__file__": path}
    with open(path) as f:
        exec(compile(f.read(), path, "exec"), scope)
    cfg = {k: v for k, v in scope.items() if not k.startswith("__") and not callable(v)
           and type(v).__name__ != "module"}
    bases = cfg.pop("_base_", [])
    if isinstance(bases, str):
        bases = [bases]
    merged = {}
    for b in bases:
        merged = _merge(merged, _load_py(os.path.normpath(os.path.join(os.path.dirname(path), b))))
    return _merge(merged, cfg)


class Config(ConfigDict):
    """``Config.fromfile(path)``: python-dict config with ``_base_`` inheritance."""

    @classmethod
    def fromfile(cls, path, import_custom_modules=True):
        raw = _load_py(os.path.abspath(path))
        cfg = cls(to_config(raw))
        if import_custom_modules and cfg.get("custom_imports"):
            ci = cfg["custom_imports"]
            for mod in ci.get("imports", []):
                try:
                    importlib.import_module(mod)
                except ImportError:
                    if not ci.get("allow_failed_imports", False):
                        raise
        return cfg

    def merge_from_dict(self, options):
        """``--cfg-options a.b.c=v`` overrides (tools/train.py:126-127)."""
        for key, v in options.items():
            d = self
            parts = key.split(".")
            for p in parts[:-1]:
                d = d.setdefault(p, ConfigDict())
            d[parts[-1]] = to_config(v)
