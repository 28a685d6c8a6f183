"""B200-native PolyphonicFormer decoder: hand-written sm_100a CUDA kernels behind a C ABI (include/pf_decoder.h,
lib/libpf_decoder.so) with a Python host mirror of the reference's mmdet-registry modules.

    from polyphonicformer_b200 import build_head, load_config
    cfg = load_config('configs/polyphonic_image/poly_r50_cityscapes_2x.py')      # the reference's config, unchanged
    roi_head = build_head(dict(cfg.model.roi_head, train_cfg=None, test_cfg=cfg.model.test_cfg.rcnn)).cuda().eval()
"""
from . import registry
from .registry import (DETECTORS, HEADS, MODELS, TRANSFORMER_LAYER, ConfigDict, Registry, build_head,
                       build_transformer_layer, load_config, register_all)
from .modules import KernelHead, KernelUpdateHead, KernelUpdateIterHead, KernelUpdator

register_all()

__all__ = ['KernelUpdator', 'KernelUpdateHead', 'KernelUpdateIterHead', 'KernelHead', 'build_head', 'build_transformer_layer',
           'load_config', 'register_all', 'Registry', 'ConfigDict', 'MODELS', 'HEADS', 'DETECTORS',
           'TRANSFORMER_LAYER', 'registry']
