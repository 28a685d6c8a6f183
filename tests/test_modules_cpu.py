"""Drop-in boundary on CPU: registry names, constructor kwargs from the reference's config, state-dict layout."""
import json
import os

import pytest
import torch

from conftest import GOLDEN
from oracle import synth

import polyphonicformer_b200 as pf

REF_CFG = '/root/reference/configs/polyphonic_image/poly_r50_cityscapes_2x.py'
ROI_HEAD_JSON = os.path.join(GOLDEN, 'roi_head_cfg.json')   # cfg.model.roi_head/test_cfg.rcnn dumped from the reference


def roi_head_cfg():
    if os.path.exists(REF_CFG):      # build container: read the reference's config file itself, unchanged
        cfg = pf.load_config(REF_CFG)
        roi, test = cfg.model.roi_head, cfg.model.test_cfg.rcnn
        dumped = json.load(open(ROI_HEAD_JSON))
        assert json.loads(json.dumps(roi)) == dumped['roi_head'], 'tests/golden/roi_head_cfg.json is stale'
        return dict(roi, train_cfg=None, test_cfg=test)
    d = json.load(open(ROI_HEAD_JSON))
    return dict(d['roi_head'], train_cfg=None, test_cfg=d['test_cfg'])


def test_registry_names():
    for name in ('KernelUpdateHead', 'KernelUpdateIterHead'):
        assert name in pf.MODELS
    assert 'KernelUpdator' in pf.TRANSFORMER_LAYER
    with pytest.raises(KeyError):
        pf.build_head(dict(type='NoSuchHead'))
    with pytest.raises(KeyError):
        pf.MODELS.register_module(module=pf.KernelUpdateHead)       # duplicate without force, like mmcv


def test_build_from_reference_config_and_state_dict_layout():
    head = pf.build_head(roi_head_cfg())
    assert isinstance(head, pf.KernelUpdateIterHead) and head.num_stages == 3 and len(head.mask_head) == 3
    assert head.mask_head[0].mask_upsample_stride == 2 and head.mask_head[-1].loss_cls.use_sigmoid
    assert head.test_cfg.max_per_img == 100 and head.num_proposals == 100
    want = {'mask_head.%d.%s' % (s, k): tuple(v) for s in range(3) for k, v in synth.stage_state_shapes().items()}
    got = {k: tuple(v.shape) for k, v in head.state_dict().items()}
    assert got == want                                     # identical keys AND shapes as the reference (strict load)
    sd = synth.synth_decoder_state(3, 0)
    res = head.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert abs(sum(p.numel() for p in head.mask_head[0].parameters()) / 1e6 - 4.02) < 0.01


def test_init_weights_matches_reference_recipe():
    head = pf.build_head(roi_head_cfg())
    head.init_weights()
    h = head.mask_head[0]
    assert torch.allclose(h.fc_cls.bias, torch.full_like(h.fc_cls.bias, -4.59511985013459))   # bias_init_with_prob(0.01)
    w = h.fc_mask.weight
    assert w.abs().max() <= (6.0 / 512) ** 0.5 + 1e-6      # xavier uniform bound for [256, 256]


def test_unsupported_configurations_fail_loudly():
    with pytest.raises(NotImplementedError):
        pf.KernelUpdator(in_channels=256, feat_channels=64, out_channels=256)
    with pytest.raises(NotImplementedError):
        pf.build_head(dict(type='KernelUpdateHead', num_classes=19, conv_kernel_size=3, num_mask_fcs=1,
                           kernel_updator_cfg=dict(type='KernelUpdator', in_channels=256, feat_channels=256,
                                                   out_channels=256)))
    head = pf.build_head(roi_head_cfg())
    x = torch.zeros(1, 256, 4, 4)
    with pytest.raises(NotImplementedError):              # no CPU / PyTorch fallback
        head.decode(x, torch.zeros(1, 111, 256, 1, 1), torch.zeros(1, 111, 4, 4), x, torch.zeros(1, 111, 256, 1, 1))
    with pytest.raises(NotImplementedError):
        head.forward_train()


def test_config_loader_merges_bases(tmp_path):
    (tmp_path / 'base.py').write_text("a = dict(x=1, y=dict(z=2, w=3))\nb = [1, 2]\n")
    (tmp_path / 'child.py').write_text("_base_ = ['./base.py']\na = dict(y=dict(z=5), q=dict(_delete_=True, k=1))\n")
    cfg = pf.load_config(str(tmp_path / 'child.py'))
    assert cfg.a.x == 1 and cfg.a.y.z == 5 and cfg.a.y.w == 3 and cfg.a.q == dict(k=1) and cfg.b == [1, 2]


RPN_HEAD_JSON = os.path.join(GOLDEN, 'rpn_head_cfg.json')   # cfg.model.rpn_head dumped from the reference's config


class _FakeNeck(torch.nn.Module):
    """Stands in for the reference's SemanticFPNWrapper (not rebuilt): accepts its config kwargs."""

    def __init__(self, **cfg):
        super().__init__()
        self.cfg = cfg
        self.dummy = torch.nn.Parameter(torch.zeros(1))

    def forward(self, img):
        return img


def rpn_head_cfg():
    if os.path.exists(REF_CFG):
        cfg = pf.load_config(REF_CFG)
        assert json.loads(json.dumps(cfg.model.rpn_head)) == json.load(open(RPN_HEAD_JSON))['rpn_head'], \
            'tests/golden/rpn_head_cfg.json is stale'
    d = json.load(open(RPN_HEAD_JSON))
    return dict(d['rpn_head'], train_cfg=None, test_cfg=d['test_cfg'])


def test_kernel_head_builds_from_reference_config():
    """`KernelHead` under the reference's name, with every kwarg of configs/_base_/models/polyphonic_former.py:30-97,
    exposes the reference's state-dict keys and refuses what the kernels do not cover."""
    assert 'KernelHead' in pf.MODELS
    real = pf.MODELS._modules['SemanticFPNWrapper']
    pf.MODELS.register_module(name='SemanticFPNWrapper', force=True, module=_FakeNeck)
    try:
        head = pf.build_head(rpn_head_cfg())
    finally:
        pf.MODELS._modules['SemanticFPNWrapper'] = real
    assert isinstance(head, pf.KernelHead) and head.num_proposals == 100
    assert head.localization_fpn.cfg['num_aux_convs'] == 2
    own = {k: tuple(v.shape) for k, v in head.state_dict().items() if not k.startswith('localization_fpn.')}
    assert own == {k: tuple(v) for k, v in synth.kernel_head_state_shapes().items()}
    missing = head.load_state_dict(synth.synth_kernel_head_state(0), strict=False)
    assert not missing.unexpected_keys and all(k.startswith('localization_fpn.') for k in missing.missing_keys)
    head.init_weights()
    with pytest.raises(NotImplementedError):
        head.eval()._decode_init_proposals([torch.zeros(1, 256, 4, 4)] * 3, [{}])      # CPU tensors: no fallback
    with pytest.raises(NotImplementedError):
        pf.KernelHead(**dict(rpn_head_cfg(), localization_fpn=_FakeNeck(), feat_refine=True))
    with pytest.raises(NotImplementedError):
        pf.KernelHead(**dict(rpn_head_cfg(), localization_fpn=_FakeNeck(), cat_stuff_mask=False))


def test_shared_feats_channel_is_identity_keyed():
    """The KernelHead -> decoder side channel matches tensor OBJECTS (weak references), never addresses: a derived
    tensor, a modified tensor, or a new tensor at a recycled address misses and takes the cast path."""
    import gc
    from polyphonicformer_b200.modules import _SharedFeats
    reg = _SharedFeats()
    x, d, feats = torch.zeros(2, 3), torch.zeros(2, 3), torch.ones(4)
    reg.publish(x, d, feats)
    assert reg.lookup(x, d) is feats
    assert reg.lookup(x.float(), d) is feats        # .float() on fp32 returns the same object ...
    assert reg.lookup(x.clone(), d) is None          # ... a copy does not
    assert reg.lookup(x[:], d) is None and reg.lookup(x, d.contiguous().view(2, 3)) is None
    x.add_(1)                                        # in-place change: version moved on
    assert reg.lookup(x, d) is None
    y = torch.zeros(2, 3)
    reg.publish(y, d, feats)
    key = id(y)
    del y
    gc.collect()
    assert key not in reg._by_id                     # entry died with the tensor


VIDEO_CFG = '/root/reference/configs/polyphonic_video/poly_r50_cityscapes_1x.py'
VIDEO_JSON = os.path.join(GOLDEN, 'video_cfg.json')   # the video-specific dicts of cfg.model, dumped from the reference's config


def video_cfg():
    d = json.load(open(VIDEO_JSON))
    if os.path.exists(VIDEO_CFG):
        m = pf.load_config(VIDEO_CFG).model
        for k in ('track_head', 'tracker', 'bbox_roi_extractor', 'track_train_cfg'):
            assert json.loads(json.dumps(m[k])) == d[k], 'tests/golden/video_cfg.json is stale (%s)' % k
    return d


class _Stub(torch.nn.Module):
    num_proposals = 100

    def __init__(self, **cfg):
        super().__init__()
        self.cfg = cfg


def test_detectors_register_and_build_without_mmdet():
    """`Polyphonic` / `PolyphonicVideo` and the tracking modules exist under the reference's names in the local registry
    (mmdet is not importable here), take the reference's kwargs, wire train_cfg / test_cfg into the heads the way
    TwoStageDetector does (two_stage.py:37-50) and expose the reference's state-dict keys."""
    from polyphonicformer_b200 import registry
    for name in ('Polyphonic', 'PolyphonicVideo', 'QuasiDenseMaskEmbedHeadGTMask', 'SingleRoIExtractor'):
        assert name in registry.MODELS, name
    assert 'QuasiDenseEmbedTracker' in registry.TRACKERS
    v = video_cfg()
    rd = json.load(open(ROI_HEAD_JSON))
    pd = json.load(open(RPN_HEAD_JSON))
    real = pf.MODELS._modules['SemanticFPNWrapper']
    pf.MODELS.register_module(name='SemanticFPNWrapper', force=True, module=_FakeNeck)
    try:
        model = registry.build_detector(dict(
            type='PolyphonicVideo', backbone=_Stub(), neck=_Stub(), rpn_head=pd['rpn_head'],
            roi_head=dict(rd['roi_head'], tracking=v['roi_head_tracking']), num_thing_classes=8, num_stuff_classes=11,
            test_cfg=dict(rpn=pd['test_cfg'], rcnn=rd['test_cfg']), track_head=v['track_head'], tracker=v['tracker'],
            bbox_roi_extractor=v['bbox_roi_extractor'], track_train_cfg=v['track_train_cfg']))
    finally:
        pf.MODELS._modules['SemanticFPNWrapper'] = real
    from polyphonicformer_b200 import detectors as d
    assert isinstance(model, d.PolyphonicVideo) and isinstance(model, d.Polyphonic)
    assert isinstance(model.rpn_head, pf.KernelHead) and isinstance(model.roi_head, pf.KernelUpdateIterHead)
    assert model.roi_head.test_cfg.max_per_img == 100 and model.roi_head.tracking is True and model.num_proposals == 100
    assert isinstance(model.track_head, d.QuasiDenseMaskEmbedHeadGTMask) and model.track_roi_extractor.num_inputs == 4
    got = {k[len('track_head.'):]: tuple(t.shape) for k, t in model.state_dict().items() if k.startswith('track_head.')}
    assert got == {k: tuple(t.shape) for k, t in synth.synth_track_head_state(0).items()}
    prefixes = {k.split('.')[0] for k in model.state_dict()}
    assert prefixes == {'rpn_head', 'roi_head', 'track_head'}, prefixes          # + backbone / neck when those have parameters
    model.init_tracker()
    assert model.cnt == 1 and isinstance(model.tracker, d.QuasiDenseEmbedTracker) and model.tracker.cfg['memo_tracklet_frames'] == 5
    with pytest.raises(NotImplementedError):
        model.forward_train(None, None)
    with pytest.raises(_cabi_error()):
        model.eval().track_head(torch.zeros(2, 256, 7, 7))                        # CPU tensor: no fallback
    with pytest.raises(NotImplementedError):
        d.QuasiDenseMaskEmbedHeadGTMask(**dict(v['track_head'], num_convs=2))
    with pytest.raises(NotImplementedError):
        d.QuasiDenseEmbedTracker(**dict({k: x for k, x in v['tracker'].items() if k != 'type'}, match_metric='cosine'))


def _cabi_error():
    from polyphonicformer_b200 import _cabi
    return _cabi.PFError


def test_semantic_fpn_wrapper_builds_from_reference_config():
    """The local `SemanticFPNWrapper` takes the reference's kwargs (configs/_base_/models/polyphonic_former.py:78-96), has the
    reference's 30 state-dict tensors and is what `KernelHead` builds when mmdet is not importable."""
    head = pf.build_head(rpn_head_cfg())
    from polyphonicformer_b200.modules import SemanticFPNWrapper, _pyramid_supported
    fpn = head.localization_fpn
    assert isinstance(fpn, SemanticFPNWrapper) and _pyramid_supported(fpn)
    got = {k: tuple(v.shape) for k, v in fpn.state_dict().items()}
    assert got == {k: tuple(v.shape) for k, v in synth.synth_semantic_fpn_state(0).items()} and len(got) == 30
    res = head.load_state_dict({**synth.synth_kernel_head_state(0),
                                **{'localization_fpn.' + k: v for k, v in synth.synth_semantic_fpn_state(0).items()}}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    with pytest.raises(NotImplementedError):
        fpn([torch.zeros(1, 256, 8, 8)] * 4)                                           # CPU tensors: no fallback
    cfg = dict(rpn_head_cfg()['localization_fpn'])
    cfg.pop('type')
    with pytest.raises(NotImplementedError):
        SemanticFPNWrapper(**dict(cfg, upsample_times=3))
    with pytest.raises(NotImplementedError):
        SemanticFPNWrapper(**dict(cfg, cat_coors=True))
