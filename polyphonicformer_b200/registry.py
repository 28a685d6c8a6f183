"""Registry / config plumbing so the reference's configs load unchanged.

The reference builds every module through mmcv ``Registry.build(cfg)`` (``mmdet/models/builder.py:1-59``,
``mmcv.cnn.bricks.transformer.TRANSFORMER_LAYER`` for ``KernelUpdator``).  mmcv is not a dependency of this package:

* when mmcv/mmdet ARE importable (a reference checkout on PYTHONPATH), ``register_all()`` registers the B200 modules
  into the reference's own registries with ``force=True`` -- that is the drop-in;
* otherwise the equally-named registries below are used.  They implement the subset of mmcv's semantics the
  reference relies on: ``register_module(name=None, force=False)`` as decorator, ``build(cfg, default_args)`` ->
  ``cls(**cfg_without_type)``, KeyError on unknown types, duplicate names rejected unless ``force``.

``load_config`` executes an mmcv-style python config (``_base_`` inheritance, recursive dict merge, ``_delete_``)
into attribute-accessible dicts.
"""
import copy
import os


class ConfigDict(dict):
    """dict with attribute access (the reference reads ``test_cfg.rcnn``, ``cfg.merge_stuff_thing`` ...)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def to_config(obj):
    if isinstance(obj, dict):
        return ConfigDict({k: to_config(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return type(obj)(to_config(v) for v in obj)
    return obj


def _merge(base, new):
    out = dict(base)
    for k, v in new.items():
        if isinstance(v, dict) and v.get('_delete_', False):
            out[k] = {kk: vv for kk, vv in v.items() if kk != '_delete_'}
        elif isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


def _load_raw(path):
    path = os.path.abspath(path)
    scope = {'__file__': path}
    with open(path) as f:
        exec(compile(f.read(), path, 'exec'), scope)
    cfg = {k: v for k, v in scope.items() if not k.startswith('__') and not callable(v)
           and type(v).__name__ != 'module'}
    bases = cfg.pop('_base_', [])
    if isinstance(bases, str):
        bases = [bases]
    merged = {}
    for b in bases:
        merged = _merge(merged, _load_raw(os.path.join(os.path.dirname(path), b)))
    return _merge(merged, cfg)


def load_config(path):
    """mmcv ``Config.fromfile`` equivalent for plain python configs."""
    return to_config(_load_raw(path))


class Registry:
    def __init__(self, name):
        self.name = name
        self._modules = {}

    def get(self, key):
        return self._modules.get(key)

    def __contains__(self, key):
        return key in self._modules

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            key = name or cls.__name__
            if key in self._modules and not force:
                raise KeyError('%s is already registered in %s' % (key, self.name))
            self._modules[key] = cls
            return cls

        if module is not None:
            return _reg(module)
        return _reg

    def build(self, cfg, default_args=None):
        if not isinstance(cfg, dict) or 'type' not in cfg:
            raise KeyError('`cfg` must be a dict with the key "type", got %r' % (cfg,))
        args = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        typ = args.pop('type')
        cls = self._modules.get(typ) if isinstance(typ, str) else typ
        if cls is None:
            raise KeyError('%s is not in the %s registry' % (typ, self.name))
        return cls(**args)


MODELS = Registry('models')          # == HEADS == DETECTORS == ROI heads (one object in mmdet too)
HEADS = DETECTORS = NECKS = MODELS
TRANSFORMER_LAYER = Registry('transformer_layer')
TRACKERS = Registry('tracker')       # polyphonic/video/qdtrack/builder.py


def build_head(cfg):
    return MODELS.build(cfg)


def build_neck(cfg):
    """The reference's own NECKS registry when a reference checkout + mmcv are importable (its SemanticFPNWrapper then keeps
    the parameters and KernelHead runs it on pf_semantic_fpn / pf_fpn_pred all the same), else this package's registry
    (modules.SemanticFPNWrapper)."""
    if cfg is None or not isinstance(cfg, dict):
        return cfg                       # an already-built module (tests, custom pipelines) or None
    try:
        from mmdet.models.builder import build_neck as mm_build_neck
        return mm_build_neck(cfg)
    except ImportError:
        return MODELS.build(cfg)


def build_transformer_layer(cfg):
    return TRANSFORMER_LAYER.build(cfg)


def build_detector(cfg, train_cfg=None, test_cfg=None):
    """mmdet.models.build_detector (builder.py:45-59) on the local registry."""
    cfg = dict(cfg)
    if train_cfg is not None:
        cfg.setdefault('train_cfg', train_cfg)
    if test_cfg is not None:
        cfg.setdefault('test_cfg', test_cfg)
    return MODELS.build(cfg)


def register_all(force=True, detectors=False):
    """Register the B200 modules under the reference's names: always into the registries above; into mmcv/mmdet's and the
    reference's own (``polyphonic.video.qdtrack.builder.TRACKERS``) when those import.  Returns whether they did.
    ``detectors``: also replace the reference's ``Polyphonic`` / ``PolyphonicVideo`` in mmdet's registry (by default the
    reference's own detector classes keep orchestrating this package's heads there)."""
    from . import detectors as d
    from . import modules as m
    heads = (m.KernelUpdateHead, m.KernelUpdateIterHead, m.KernelHead, d.QuasiDenseMaskEmbedHeadGTMask)
    for cls in heads + (d.SingleRoIExtractor, d.Polyphonic, d.PolyphonicVideo, m.SemanticFPNWrapper):
        MODELS.register_module(force=True, module=cls)
    TRANSFORMER_LAYER.register_module(force=True, module=m.KernelUpdator)
    TRACKERS.register_module(force=True, module=d.QuasiDenseEmbedTracker)
    try:   # the reference's own registries (drop-in when a reference checkout + mmcv are installed)
        from mmdet.models.builder import HEADS as MM_HEADS
        from mmcv.cnn.bricks.transformer import TRANSFORMER_LAYER as MM_TL
    except Exception:
        return False
    for cls in heads:
        MM_HEADS.register_module(force=force, module=cls)
    MM_TL.register_module(force=force, module=m.KernelUpdator)
    try:
        from polyphonic.video.qdtrack.builder import TRACKERS as REF_TRACKERS
        REF_TRACKERS.register_module(force=force, module=d.QuasiDenseEmbedTracker)
    except Exception:
        pass
    if detectors:
        from mmdet.models.builder import DETECTORS as MM_DET
        MM_DET.register_module(force=force, module=d.Polyphonic)
        MM_DET.register_module(force=force, module=d.PolyphonicVideo)
    return True
