"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

A minimal stand-in for the reference's un-vendored dependency ``mmcv`` (pinned
1.3.18 at /root/reference/scripts/docker_env/Dockerfile:11-12) plus empty stubs
for ``pycocotools``, ``terminaltables``, ``panopticapi``, ``matplotlib`` and
``timm``, so that the reference's own, unmodified ``polyphonic/`` and ``mmdet/``
packages can be imported in the build container::

    import oracle.mmcv_shim as shim; shim.install(); import polyphonic

It exists for exactly one purpose: run the reference's real forward on CPU to
generate the golden vectors under ``tests/golden/`` (``oracle/make_golden.py``)
that pin ``oracle/decoder_ref.py``.  It cannot travel to the GPU box
(/root/reference is absent there) and nothing at run time depends on it.

Only the pieces of mmcv that carry arithmetic or registry semantics on the
decoder path are restated (SURVEY.md section 8c lists the call sites):

* ``Registry`` / ``build_from_cfg``            (mmcv/utils/registry.py semantics)
* ``ConvModule``: conv -> norm -> act, ``bias='auto'`` means bias iff no norm,
  sub-module names ``conv`` / ``gn`` / ``bn`` / ``activate``
* ``build_norm_layer``: LN eps 1e-5, GN eps 1e-5, returned name ``ln`` / ``gn``
* ``MultiheadAttention``: ``identity + nn.MultiheadAttention(q, k, v)[0]``,
  sequence-first, third positional ctor arg = attn_drop
* ``FFN``: ``identity + Sequential(Sequential(Linear, act, Dropout), Linear,
  Dropout)(x)``; legacy kwarg ``dropout`` maps to ``ffn_drop``
* ``RoIAlign`` -> torchvision ``roi_align(..., aligned=True)``
Everything else under the stub roots is fabricated on demand as inert objects.
"""
import copy
import importlib.abc
import importlib.machinery
import inspect
import logging
import math
import sys
import types

import torch
import torch.nn as nn

STUB_ROOTS = ('mmcv', 'pycocotools', 'terminaltables', 'panopticapi',
              'matplotlib', 'timm')


# --------------------------------------------------------------------------
# registry / config plumbing
# --------------------------------------------------------------------------
class AttrDict(dict):
    """dict with attribute access (stands in for mmcv ConfigDict)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return AttrDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def to_attr(obj):
    if isinstance(obj, dict):
        return AttrDict({k: to_attr(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return type(obj)(to_attr(v) for v in obj)
    return obj


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict):
        raise TypeError(f'cfg must be a dict, got {type(cfg)}')
    args = dict(cfg)
    if default_args is not None:
        for k, v in default_args.items():
            args.setdefault(k, v)
    obj_type = args.pop('type')
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError(f'{obj_type} is not in the {registry.name} registry')
    else:
        obj_cls = obj_type
    return obj_cls(**args)


class Registry:

    def __init__(self, name, build_func=None, parent=None, scope=None):
        self._name = name
        self._module_dict = {}
        self._children = {}
        self.parent = parent
        if build_func is None:
            build_func = parent.build_func if parent is not None else build_from_cfg
        self.build_func = build_func
        if parent is not None:
            parent._children[name] = self

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def __contains__(self, key):
        return self.get(key) is not None

    def __len__(self):
        return len(self._module_dict)

    def get(self, key):
        if key in self._module_dict:
            return self._module_dict[key]
        if self.parent is not None:
            return self.parent.get(key)
        return None

    def build(self, *args, **kwargs):
        return self.build_func(*args, **kwargs, registry=self)

    def _register_module(self, module_class, module_name=None, force=False):
        if module_name is None:
            module_name = module_class.__name__
        names = [module_name] if isinstance(module_name, str) else module_name
        for n in names:
            if not force and n in self._module_dict:
                raise KeyError(f'{n} is already registered in {self.name}')
            self._module_dict[n] = module_class

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register_module(module, name, force)
            return module

        def _register(cls):
            self._register_module(cls, name, force)
            return cls
        return _register


def build_model_from_cfg(cfg, registry, default_args=None):
    if isinstance(cfg, list):
        return nn.Sequential(*[build_from_cfg(c, registry, default_args) for c in cfg])
    return build_from_cfg(cfg, registry, default_args)


def digit_version(version_str, length=4):
    out = []
    for x in str(version_str).split('+')[0].split('.'):
        if x.isdigit():
            out.append(int(x))
        elif 'rc' in x:
            a, b = x.split('rc')
            out += [int(a) - 1, int(b)]
    return tuple(out)


# --------------------------------------------------------------------------
# cnn bricks that carry arithmetic
# --------------------------------------------------------------------------
MODELS = Registry('model', build_func=build_model_from_cfg)
CONV_LAYERS = Registry('conv layer')
NORM_LAYERS = Registry('norm layer')
ACTIVATION_LAYERS = Registry('activation layer')
PADDING_LAYERS = Registry('padding layer')
UPSAMPLE_LAYERS = Registry('upsample layer')
PLUGIN_LAYERS = Registry('plugin layer')
DROPOUT_LAYERS = Registry('drop out layers')
POSITIONAL_ENCODING = Registry('position encoding')
ATTENTION = Registry('attention')
FEEDFORWARD_NETWORK = Registry('feed-forward Network')
TRANSFORMER_LAYER = Registry('transformerLayer')
TRANSFORMER_LAYER_SEQUENCE = Registry('transformer-layers sequence')

CONV_LAYERS.register_module('Conv2d', module=nn.Conv2d)
CONV_LAYERS.register_module('Conv', module=nn.Conv2d)
for _n, _m in dict(ReLU=nn.ReLU, LeakyReLU=nn.LeakyReLU, PReLU=nn.PReLU, ReLU6=nn.ReLU6,
                   ELU=nn.ELU, Sigmoid=nn.Sigmoid, Tanh=nn.Tanh, GELU=nn.GELU).items():
    ACTIVATION_LAYERS.register_module(_n, module=_m)
DROPOUT_LAYERS.register_module('Dropout', module=nn.Dropout)


def build_conv_layer(cfg, *args, **kwargs):
    cfg = dict(type='Conv2d') if cfg is None else dict(cfg)
    layer = CONV_LAYERS.get(cfg.pop('type'))
    return layer(*args, **kwargs, **cfg)


def build_norm_layer(cfg, num_features, postfix=''):
    cfg = dict(cfg)
    t = cfg.pop('type')
    requires_grad = cfg.pop('requires_grad', True)
    cfg.setdefault('eps', 1e-5)
    if t in ('BN', 'BN2d', 'SyncBN'):
        layer, abbr = nn.BatchNorm2d(num_features, **cfg), 'bn'
    elif t == 'GN':
        layer, abbr = nn.GroupNorm(num_channels=num_features, **cfg), 'gn'
    elif t == 'LN':
        layer, abbr = nn.LayerNorm(num_features, **cfg), 'ln'
    else:
        raise KeyError(t)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return abbr + str(postfix), layer


def build_activation_layer(cfg):
    return build_from_cfg(cfg, ACTIVATION_LAYERS)


def build_dropout(cfg, default_args=None):
    return build_from_cfg(cfg, DROPOUT_LAYERS, default_args)


def build_plugin_layer(cfg, postfix='', **kwargs):
    raise NotImplementedError('plugins are not used by PolyphonicFormer')


def bias_init_with_prob(prior_prob):
    return float(-math.log((1 - prior_prob) / prior_prob))


def constant_init(module, val, bias=0):
    if getattr(module, 'weight', None) is not None:
        nn.init.constant_(module.weight, val)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


def normal_init(module, mean=0, std=1, bias=0):
    if getattr(module, 'weight', None) is not None:
        nn.init.normal_(module.weight, mean, std)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


def xavier_init(module, gain=1, bias=0, distribution='normal'):
    if getattr(module, 'weight', None) is not None:
        if distribution == 'uniform':
            nn.init.xavier_uniform_(module.weight, gain=gain)
        else:
            nn.init.xavier_normal_(module.weight, gain=gain)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


def kaiming_init(module, a=0, mode='fan_out', nonlinearity='relu', bias=0, distribution='normal'):
    if getattr(module, 'weight', None) is not None:
        if distribution == 'uniform':
            nn.init.kaiming_uniform_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
        else:
            nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


class BaseModule(nn.Module):

    def __init__(self, init_cfg=None):
        super().__init__()
        self._is_init = False
        self.init_cfg = copy.deepcopy(init_cfg)

    @property
    def is_init(self):
        return self._is_init

    def init_weights(self):
        for m in self.children():
            if hasattr(m, 'init_weights'):
                m.init_weights()


class Sequential(BaseModule, nn.Sequential):

    def __init__(self, *args, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.Sequential.__init__(self, *args)


class ModuleList(BaseModule, nn.ModuleList):

    def __init__(self, modules=None, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.ModuleList.__init__(self, modules)


class ConvModule(nn.Module):
    """conv -> norm -> act; bias='auto' => bias iff norm_cfg is None."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0,
                 dilation=1, groups=1, bias='auto', conv_cfg=None, norm_cfg=None,
                 act_cfg=dict(type='ReLU'), inplace=True, with_spectral_norm=False,
                 padding_mode='zeros', order=('conv', 'norm', 'act')):
        super().__init__()
        assert order == ('conv', 'norm', 'act') and not with_spectral_norm
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == 'auto':
            bias = not self.with_norm
        self.conv = build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size,
                                     stride=stride, padding=padding, dilation=dilation,
                                     groups=groups, bias=bias)
        self.norm_name = None
        if self.with_norm:
            self.norm_name, norm = build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        if self.with_activation:
            act_cfg = dict(act_cfg)
            if act_cfg['type'] not in ('Tanh', 'PReLU', 'Sigmoid', 'GELU'):
                act_cfg.setdefault('inplace', inplace)
            self.activate = build_activation_layer(act_cfg)
        nonlin = 'relu'
        kaiming_init(self.conv, a=0, nonlinearity=nonlin)
        if self.with_norm:
            constant_init(getattr(self, self.norm_name), 1, bias=0)

    @property
    def norm(self):
        return getattr(self, self.norm_name) if self.norm_name else None

    def forward(self, x, activate=True, norm=True):
        x = self.conv(x)
        if norm and self.with_norm:
            x = self.norm(x)
        if activate and self.with_activation:
            x = self.activate(x)
        return x


class MultiheadAttention(BaseModule):
    """mmcv 1.3.18 wrapper: out = identity + dropout(proj_drop(attn(q,k,v)[0]))."""

    def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0.,
                 dropout_layer=dict(type='Dropout', drop_prob=0.), init_cfg=None,
                 batch_first=False, **kwargs):
        super().__init__(init_cfg)
        if 'dropout' in kwargs:
            attn_drop = kwargs.pop('dropout')
        self.embed_dims, self.num_heads, self.batch_first = embed_dims, num_heads, batch_first
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop, **kwargs)
        self.proj_drop = nn.Dropout(proj_drop)
        self.dropout_layer = nn.Identity()

    def forward(self, query, key=None, value=None, identity=None, query_pos=None,
                key_pos=None, attn_mask=None, key_padding_mask=None, **kwargs):
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
            key_pos = query_pos
        if query_pos is not None:
            query = query + query_pos
        if key_pos is not None:
            key = key + key_pos
        if self.batch_first:
            query, key, value = (t.transpose(0, 1) for t in (query, key, value))
        out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask,
                        key_padding_mask=key_padding_mask)[0]
        if self.batch_first:
            out = out.transpose(0, 1)
        return identity + self.dropout_layer(self.proj_drop(out))


class FFN(BaseModule):
    """mmcv 1.3.18 FFN: out = identity + layers(x)."""

    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2,
                 act_cfg=dict(type='ReLU', inplace=True), ffn_drop=0., dropout_layer=None,
                 add_identity=True, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        if 'dropout' in kwargs:
            ffn_drop = kwargs.pop('dropout')
        assert num_fcs >= 2
        self.embed_dims, self.feedforward_channels, self.num_fcs = embed_dims, feedforward_channels, num_fcs
        self.activate = build_activation_layer(act_cfg)
        layers, c = [], embed_dims
        for _ in range(num_fcs - 1):
            layers.append(Sequential(nn.Linear(c, feedforward_channels), self.activate,
                                     nn.Dropout(ffn_drop)))
            c = feedforward_channels
        layers.append(nn.Linear(feedforward_channels, embed_dims))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = Sequential(*layers)
        self.dropout_layer = nn.Identity()
        self.add_identity = add_identity

    def forward(self, x, identity=None):
        out = self.layers(x)
        if not self.add_identity:
            return self.dropout_layer(out)
        if identity is None:
            identity = x
        return identity + self.dropout_layer(out)


def build_transformer_layer(cfg, default_args=None):
    return build_from_cfg(cfg, TRANSFORMER_LAYER, default_args)


def build_positional_encoding(cfg, default_args=None):
    return build_from_cfg(cfg, POSITIONAL_ENCODING, default_args)


def build_attention(cfg, default_args=None):
    return build_from_cfg(cfg, ATTENTION, default_args)


def build_feedforward_network(cfg, default_args=None):
    return build_from_cfg(cfg, FEEDFORWARD_NETWORK, default_args)


class RoIAlign(nn.Module):

    def __init__(self, output_size, spatial_scale=1.0, sampling_ratio=0, pool_mode='avg',
                 aligned=True, use_torchvision=False):
        super().__init__()
        self.output_size = (output_size, output_size) if isinstance(output_size, int) else tuple(output_size)
        self.spatial_scale, self.sampling_ratio, self.aligned = float(spatial_scale), int(sampling_ratio), aligned
        assert pool_mode == 'avg'

    def forward(self, x, rois):
        from torchvision.ops import roi_align
        return roi_align(x, rois, self.output_size, self.spatial_scale, self.sampling_ratio, self.aligned)


def _identity_decorator_factory(*dargs, **dkwargs):
    """force_fp32 / auto_fp16: usable bare or with arguments; no-op."""
    if len(dargs) == 1 and callable(dargs[0]) and not dkwargs:
        return dargs[0]
    return lambda f: f


def get_logger(name, log_file=None, log_level=logging.INFO, file_mode='w'):
    return logging.getLogger(name)


def print_log(msg, logger=None, level=logging.INFO):
    pass


# --------------------------------------------------------------------------
# auto-stub import hook
# --------------------------------------------------------------------------
class _Anything:
    """Inert object usable as function, decorator (with or without args), iterable."""

    def __call__(self, *a, **k):
        if len(a) == 1 and not k and (inspect.isfunction(a[0]) or inspect.isclass(a[0])):
            return a[0]
        return self

    def __getattr__(self, n):
        if n.startswith('__'):
            raise AttributeError(n)
        return _Anything()

    def __iter__(self):
        return iter(())


class _StubRegistry(Registry):
    def get(self, key):
        got = super().get(key)
        return got

    def _register_module(self, module_class, module_name=None, force=False):
        super()._register_module(module_class, module_name, True)


_REGISTRY_NAMES = {'HOOKS', 'MODELS', 'ATTENTION', 'PIPELINES', 'DATASETS', 'RUNNERS', 'OPTIMIZERS'}


def _fabricate(modname, attr):
    if attr.startswith('__'):
        raise AttributeError(attr)
    if attr.isupper() and ('_' in attr or attr in _REGISTRY_NAMES):
        return _StubRegistry(f'{modname}.{attr}')
    if attr[0].isupper():
        return type(attr, (nn.Module,), {'__module__': modname, '_is_stub': True})
    return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, attr):
        val = _fabricate(self.__name__, attr)
        setattr(self, attr, val)
        return val


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split('.')[0] in STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        _populate(module)


def _populate(m):
    n = m.__name__
    if n == 'mmcv':
        m.__version__ = '1.3.18'
        m.ConfigDict = AttrDict
        m.Config = types.SimpleNamespace(fromfile=load_config)
        m.is_str = lambda x: isinstance(x, str)
        m.is_list_of = lambda seq, t: isinstance(seq, list) and all(isinstance(i, t) for i in seq)
        m.is_tuple_of = lambda seq, t: isinstance(seq, tuple) and all(isinstance(i, t) for i in seq)
        m.is_seq_of = lambda seq, t, seq_type=None: all(isinstance(i, t) for i in seq)
    elif n == 'mmcv.utils':
        m.Registry, m.build_from_cfg, m.digit_version = Registry, build_from_cfg, digit_version
        m.TORCH_VERSION = torch.__version__
        m.get_logger, m.print_log = get_logger, print_log
        m.ConfigDict = AttrDict
        m.to_2tuple = lambda x: x if isinstance(x, (tuple, list)) else (x, x)
        m.is_str = lambda x: isinstance(x, str)
    elif n == 'mmcv.cnn':
        for k in ('MODELS', 'CONV_LAYERS', 'NORM_LAYERS', 'ACTIVATION_LAYERS', 'PLUGIN_LAYERS',
                  'UPSAMPLE_LAYERS', 'PADDING_LAYERS', 'ConvModule', 'build_conv_layer',
                  'build_norm_layer', 'build_activation_layer', 'build_plugin_layer',
                  'bias_init_with_prob', 'normal_init', 'constant_init', 'kaiming_init',
                  'xavier_init', 'build_model_from_cfg'):
            setattr(m, k, globals()[k])
        m.Linear, m.Conv2d = nn.Linear, nn.Conv2d
    elif n in ('mmcv.cnn.bricks.transformer', 'mmcv.cnn.bricks.registry', 'mmcv.cnn.bricks'):
        for k in ('TRANSFORMER_LAYER', 'TRANSFORMER_LAYER_SEQUENCE', 'POSITIONAL_ENCODING',
                  'ATTENTION', 'FEEDFORWARD_NETWORK', 'CONV_LAYERS', 'NORM_LAYERS',
                  'ACTIVATION_LAYERS', 'PLUGIN_LAYERS', 'DROPOUT_LAYERS', 'PADDING_LAYERS',
                  'UPSAMPLE_LAYERS', 'build_transformer_layer', 'build_positional_encoding',
                  'build_attention', 'build_feedforward_network', 'build_dropout',
                  'MultiheadAttention', 'FFN', 'ConvModule', 'build_norm_layer',
                  'build_activation_layer', 'build_conv_layer'):
            setattr(m, k, globals()[k])
    elif n in ('mmcv.runner', 'mmcv.runner.base_module'):
        m.BaseModule, m.ModuleList, m.Sequential = BaseModule, ModuleList, Sequential
        m.force_fp32 = m.auto_fp16 = _identity_decorator_factory
        m.get_dist_info = lambda: (0, 1)
    elif n == 'mmcv.ops':
        m.RoIAlign = RoIAlign


# --------------------------------------------------------------------------
# config loader (python configs with _base_ inheritance)
# --------------------------------------------------------------------------
def _merge(base, new):
    out = dict(base)
    for k, v in new.items():
        if isinstance(v, dict) and v.get('_delete_', False):
            v = {kk: vv for kk, vv in v.items() if kk != '_delete_'}
            out[k] = v
        elif isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


def _load_raw(path):
    import os
    ns = {}
    with open(path) as f:
        exec(compile(f.read(), path, 'exec'), ns)
    cfg = {k: v for k, v in ns.items()
           if not k.startswith('__') and not isinstance(v, types.ModuleType) and not callable(v)}
    bases = cfg.pop('_base_', [])
    if isinstance(bases, str):
        bases = [bases]
    merged = {}
    for b in bases:
        merged = _merge(merged, _load_raw(os.path.join(os.path.dirname(path), b)))
    return _merge(merged, cfg)


def load_config(path):
    return to_attr(_load_raw(path))


_installed = False


def install(reference_root='/root/reference'):
    """Install the stub finder and put the reference tree on sys.path."""
    global _installed
    if _installed:
        return
    sys.meta_path.insert(0, _StubFinder())
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    _installed = True
