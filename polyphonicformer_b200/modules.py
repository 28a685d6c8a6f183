"""Drop-in modules for the decoder hot path, registered under the reference's names.

``KernelUpdator``, ``KernelUpdateHead``, ``KernelUpdateIterHead`` and ``KernelHead`` take the reference's constructor kwargs
(``configs/_base_/models/polyphonic_former.py:99-164``), expose the reference's ``state_dict`` keys and shapes
 so its checkpoints load with ``strict=True``, and keep the reference's forward signatures:

  KernelUpdator.forward(update_feature, input_feature)           polyphonic/funcs/kernel_updator.py:55-93
  KernelUpdateHead.forward(x, proposal_feat, mask_preds, ...)    polyphonic/kernel_update_head.py:212-353
  KernelUpdateIterHead._mask_forward / simple_test_mask_preds / simple_test
                                                                  polyphonic/kernel_update.py:125-157, 282-401
  KernelHead._decode_init_proposals / simple_test_rpn             polyphonic/kernel_head.py:240-347, 700-706

The ``nn`` layers below only HOLD parameters.  All forward arithmetic runs in libpf_decoder.so (sm_100a CUDA) through
``DecoderEngine``; there is no PyTorch fallback -- on a non-CUDA tensor, or an unsupported configuration, the modules
raise.  Inference only (``forward_train`` / losses are out of scope).
"""
import math
import weakref

import torch
import torch.nn as nn

from . import _cabi
from .decoder import DecoderEngine
from .registry import ConfigDict, build_head, build_neck, build_transformer_layer, to_config

__all__ = ['KernelUpdator', 'KernelUpdateHead', 'KernelUpdateIterHead', 'KernelHead']


def _unsupported(what):
    raise NotImplementedError('polyphonicformer_b200: %s is not supported by the sm_100a decoder kernels '
                              '(only the configuration family shipped in configs/_base_/models/polyphonic_former.py); '
                              'there is no PyTorch fallback' % what)


def _version_key(module):
    return tuple((p.data_ptr(), p._version) for p in module.parameters())


class _SharedFeats:
    """Explicit side channel between ``KernelHead`` (producer) and ``KernelUpdateIterHead`` / ``KernelUpdateHead``
    (consumers): the feature maps ``KernelHead`` returns are views of (or fp32 copies next to) the decoder's bf16 feature
    buffer ``[2][B][256][HWp]``; this table maps the IDENTITY of those two returned tensor objects to that buffer, so a
    consumer that receives exactly those objects skips the re-cast.  Entries hold weak references only: when the
    returned tensors die the entry dies with them, and a tensor that merely reuses the address (the next frame through
    the caching allocator), or any derived tensor (``.float()``, ``.contiguous()``, a slice), never matches -- it takes
    the cast path, which is correct for any input."""

    def __init__(self):
        self._by_id = {}

    def publish(self, x, depth, feats):
        key = id(x)
        self._by_id[key] = (weakref.ref(x, lambda _r, k=key: self._by_id.pop(k, None)), weakref.ref(depth), feats,
                            x._version, depth._version)

    def lookup(self, x, depth):
        e = self._by_id.get(id(x))
        if e is None or e[0]() is not x or e[1]() is not depth or e[3] != x._version or e[4] != depth._version:
            return None
        return e[2]


SHARED_FEATS = _SharedFeats()


class KernelUpdator(nn.Module):
    """Adaptive kernel update (reference: polyphonic/funcs/kernel_updator.py:6-93)."""

    def __init__(self, in_channels=256, feat_channels=64, out_channels=None, input_feat_shape=3, gate_sigmoid=True,
                 gate_norm_act=False, activate_out=False, act_cfg=dict(type='ReLU', inplace=True),
                 norm_cfg=dict(type='LN')):
        super().__init__()
        self.in_channels = in_channels
        self.feat_channels = feat_channels
        self.out_channels_raw = out_channels
        self.gate_sigmoid = gate_sigmoid
        self.gate_norm_act = gate_norm_act
        self.activate_out = activate_out
        if isinstance(input_feat_shape, int):
            input_feat_shape = [input_feat_shape] * 2
        self.input_feat_shape = input_feat_shape
        self.act_cfg = act_cfg
        self.norm_cfg = norm_cfg
        self.out_channels = out_channels if out_channels else in_channels
        if not (in_channels == feat_channels == self.out_channels == _cabi.PF_C):
            _unsupported('KernelUpdator with in/feat/out channels != 256')
        if not gate_sigmoid or gate_norm_act or activate_out:
            _unsupported('KernelUpdator(gate_sigmoid=False | gate_norm_act=True | activate_out=True)')
        if norm_cfg.get('type') != 'LN' or act_cfg.get('type') != 'ReLU':
            _unsupported('KernelUpdator norm/activation other than LN/ReLU')
        self.num_params_in = self.feat_channels
        self.num_params_out = self.feat_channels
        self.dynamic_layer = nn.Linear(in_channels, self.num_params_in + self.num_params_out)
        self.input_layer = nn.Linear(in_channels, self.num_params_in + self.num_params_out, 1)
        self.input_gate = nn.Linear(in_channels, feat_channels, 1)
        self.update_gate = nn.Linear(in_channels, feat_channels, 1)
        self.norm_in = nn.LayerNorm(feat_channels)
        self.norm_out = nn.LayerNorm(feat_channels)
        self.input_norm_in = nn.LayerNorm(feat_channels)
        self.input_norm_out = nn.LayerNorm(feat_channels)
        self.activation = nn.ReLU(inplace=True)
        self.fc_layer = nn.Linear(feat_channels, self.out_channels, 1)
        self.fc_norm = nn.LayerNorm(self.out_channels)
        self._packed = None

    def forward(self, update_feature, input_feature):
        """update_feature [..., 256] (R rows), input_feature [R, 1, 256] -> [R, 1, 256]."""
        from .decoder import PackedUpdator, run_kernel_updator
        if not update_feature.is_cuda:
            _unsupported('KernelUpdator.forward on a %s tensor' % update_feature.device.type)
        key = _version_key(self)
        if self._packed is None or self._packed[0] != key:
            self._packed = (key, PackedUpdator(self.state_dict(), update_feature.device))
        upd = update_feature.reshape(-1, self.in_channels).float().contiguous()
        R = upd.shape[0]
        inp = input_feature.reshape(R, -1, self.feat_channels)
        if inp.shape[1] != 1:
            _unsupported('conv_kernel_size != 1 (K*K=%d kernel positions)' % inp.shape[1])
        out = run_kernel_updator(self._packed[1], upd, inp.reshape(R, self.feat_channels).float().contiguous())
        return out.reshape(R, 1, self.out_channels)


class _MHA(nn.Module):
    """Parameter holder with mmcv MultiheadAttention's layout: ``attn`` = nn.MultiheadAttention."""

    def __init__(self, embed_dims, num_heads, dropout):
        super().__init__()
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, dropout)


class _FFN(nn.Module):
    """Parameter holder with mmcv FFN's layout: layers.0.0 = Linear(C, F), layers.1 = Linear(F, C)."""

    def __init__(self, embed_dims, feedforward_channels, dropout):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True), nn.Dropout(dropout)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(dropout))


class _Conv1x1(nn.Module):
    """Parameter holder with mmcv ConvModule's layout (``conv``), bias because there is no norm."""

    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 1)


class _LossCfg(ConfigDict):
    """Inference only needs ``loss_cls.use_sigmoid`` (kernel_update.py:333)."""


class KernelUpdateHead(nn.Module):
    """One decoder stage (reference: polyphonic/kernel_update_head.py:18-353).

    The constructor DEFAULTS are the reference's (num_mask_fcs=3, conv_kernel_size=3, a 'DynamicConv' updator with 64
    feature channels) so that configs which rely on them mean the same thing here -- but the sm_100a kernels implement
    the configuration the shipped configs select (configs/_base_/models/polyphonic_former.py:111-164: num_mask_fcs=1,
    conv_kernel_size=1, kernel_updator_cfg type 'KernelUpdator' with 256 channels), and anything else raises in the
    constructor: ``KernelUpdateHead()`` with no arguments is therefore an error, as documented in INTEGRATION.md."""

    def __init__(self, num_classes=80, num_thing_classes=80, num_stuff_classes=53, num_ffn_fcs=2, num_heads=8,
                 num_cls_fcs=1, num_mask_fcs=3, feedforward_channels=2048, in_channels=256, out_channels=256,
                 dropout=0.0, mask_thr=0.5, act_cfg=dict(type='ReLU', inplace=True),
                 ffn_act_cfg=dict(type='ReLU', inplace=True), conv_kernel_size=3, feat_transform_cfg=None,
                 hard_mask_thr=0.5, kernel_init=False, with_ffn=True, mask_out_stride=4, relative_coors=False,
                 relative_coors_off=False, feat_gather_stride=1, mask_transform_stride=1, mask_upsample_stride=1,
                 mask_assign_stride=4, ignore_label=255,
                 kernel_updator_cfg=dict(type='DynamicConv', in_channels=256, feat_channels=64, out_channels=256,
                                         input_feat_shape=1, act_cfg=dict(type='ReLU', inplace=True),
                                         norm_cfg=dict(type='LN')),
                 loss_rank=None, loss_mask=dict(type='CrossEntropyLoss', use_mask=True, loss_weight=1.0),
                 loss_dice=dict(type='DiceLoss', loss_weight=3.0),
                 loss_cls=dict(type='FocalLoss', use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=2.0),
                 loss_depth=dict(type='DepthLoss', loss_weight=1.0, act=True, si_weight=1.0, sq_rel_weight=1.0,
                                 abs_rel_weight=1.0),
                 depth_act_mode='monodepth'):
        super().__init__()
        self.num_classes = num_classes
        self.loss_cls = _LossCfg(to_config(loss_cls))
        self.loss_cls.setdefault('use_sigmoid', False)
        self.loss_mask, self.loss_dice = to_config(loss_mask), to_config(loss_dice)
        self.loss_depth, self.loss_rank = to_config(loss_depth), to_config(loss_rank)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.mask_thr = mask_thr
        self.fp16_enabled = False
        self.dropout = dropout
        self.num_heads = num_heads
        self.hard_mask_thr = hard_mask_thr
        self.kernel_init = kernel_init
        self.with_ffn = with_ffn
        self.mask_out_stride = mask_out_stride
        self.relative_coors, self.relative_coors_off = relative_coors, relative_coors_off
        self.conv_kernel_size = conv_kernel_size
        self.feat_gather_stride = feat_gather_stride
        self.mask_transform_stride = mask_transform_stride
        self.mask_upsample_stride = mask_upsample_stride
        self.num_thing_classes, self.num_stuff_classes = num_thing_classes, num_stuff_classes
        self.mask_assign_stride = mask_assign_stride
        self.ignore_label = ignore_label
        self.feedforward_channels = feedforward_channels
        self.depth_act_mode = depth_act_mode

        if conv_kernel_size != 1:
            _unsupported('conv_kernel_size=%d' % conv_kernel_size)
        if in_channels != _cabi.PF_C or out_channels != _cabi.PF_C or num_heads != _cabi.PF_HEADS:
            _unsupported('in/out channels != 256 or num_heads != 8')
        if not with_ffn or num_ffn_fcs != 2 or num_cls_fcs != 1 or num_mask_fcs != 1:
            _unsupported('with_ffn=False / num_ffn_fcs != 2 / num_cls_fcs != 1 / num_mask_fcs != 1')
        if feat_gather_stride != 1 or mask_transform_stride != 1 or hard_mask_thr != 0.5 or dropout != 0.0:
            _unsupported('feat_gather_stride / mask_transform_stride != 1, hard_mask_thr != 0.5 or dropout > 0')
        if not self.loss_cls.use_sigmoid:
            _unsupported('softmax classification (loss_cls.use_sigmoid=False)')
        if num_classes > _cabi.PF_MAX_CLASSES or feedforward_channels % 256:
            _unsupported('num_classes > 32 or feedforward_channels not a multiple of 256')

        C = in_channels
        self.attention = _MHA(C, num_heads, dropout)
        self.attention_depth = _MHA(C, num_heads, dropout)
        self.attention_norm = nn.LayerNorm(C)
        self.attention_norm_depth = nn.LayerNorm(C)
        self.kernel_update_conv = build_transformer_layer(kernel_updator_cfg)
        self.kernel_update_conv_depth = build_transformer_layer(kernel_updator_cfg)
        if feat_transform_cfg is not None:
            kernel_size = feat_transform_cfg.pop('kernel_size', 1)   # the reference mutates the shared dict too (:125)
            if kernel_size != 1 or feat_transform_cfg.get('act_cfg', None) is not None \
                    or feat_transform_cfg.get('norm_cfg', None) is not None:
                _unsupported('feat_transform other than a bare 1x1 conv')
            self.feat_transform = _Conv1x1(C)
            self.feat_depth_transform = _Conv1x1(C)
        else:
            self.feat_transform = None
            self.feat_depth_transform = None
        self.ffn = _FFN(C, feedforward_channels, dropout)
        self.ffn_norm = nn.LayerNorm(C)
        self.ffn_depth = _FFN(C, feedforward_channels, dropout)
        self.ffn_norm_depth = nn.LayerNorm(C)
        self.cls_fcs = nn.ModuleList([nn.Linear(C, C, bias=False), nn.LayerNorm(C), nn.ReLU(inplace=True)])
        self.fc_cls = nn.Linear(C, num_classes)
        self.mask_fcs = nn.ModuleList([nn.Linear(C, C, bias=False), nn.LayerNorm(C), nn.ReLU(inplace=True)])
        self.depth_regs = nn.ModuleList([nn.Linear(C, C, bias=False), nn.LayerNorm(C)])
        self.fc_mask = nn.Linear(C, out_channels)
        self.fc_depth = nn.Linear(C, out_channels)
        self._engine = None
        self._feats_cache = None

    def init_weights(self):
        """kernel_update_head.py:193-210: xavier-uniform matrices, focal-loss prior on fc_cls.bias."""
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        if self.loss_cls.use_sigmoid:
            nn.init.constant_(self.fc_cls.bias, float(-math.log((1 - 0.01) / 0.01)))
        if self.kernel_init:
            nn.init.normal_(self.fc_mask.weight, mean=0, std=0.01)

    # ------------------------------------------------------------------ engine plumbing
    def engine(self, device):
        key = (_version_key(self), str(device))
        if self._engine is None or self._engine[0] != key:
            self._engine = (key, DecoderEngine([self.state_dict()], device, self.num_classes,
                                               self.feedforward_channels))
        return self._engine[1]

    def _prepared_feats(self, x, depth_feats):
        """The decoder-layout bf16 copy of (x, depth_feats).  Reused only for the very same tensor OBJECTS at the same
        version: via SHARED_FEATS when KernelHead produced them, else via a one-entry cache that holds weak references
        (an address recycled by the caching allocator for the next frame is a different object and misses)."""
        shared = SHARED_FEATS.lookup(x, depth_feats)
        if shared is not None:
            return shared
        c = self._feats_cache
        if c is not None and c[0]() is x and c[1]() is depth_feats and c[2] == (x._version, depth_feats._version):
            return c[3]
        feats = self.engine(x.device).prepare_feats(x, depth_feats)
        self._feats_cache = (weakref.ref(x), weakref.ref(depth_feats), (x._version, depth_feats._version), feats)
        return feats

    def forward(self, x, proposal_feat, mask_preds, prev_cls_score=None, mask_shape=None, img_metas=None,
                depth_preds=None, depth_proposal=None, depth_feats=None, _feats=None):
        """Returns (cls_score [B,N,classes], new_mask_preds [B,N,H,W], obj_feat [B,N,C,1,1],
        new_depth_preds [B,N,H,W], depth_feat_new [B,N,C,1,1]) like the reference."""
        if not x.is_cuda:
            _unsupported('KernelUpdateHead.forward on a %s tensor' % x.device.type)
        B, N = proposal_feat.shape[:2]
        C, H, W = x.shape[-3:]
        if mask_preds.shape[-2:] != (H, W):
            _unsupported('mask_preds at a different resolution than the feature map')
        if mask_shape is not None and mask_shape[0] != H:
            _unsupported('mask_shape resizing')
        eng = self.engine(x.device)
        feats = _feats if _feats is not None else self._prepared_feats(x, depth_feats)
        cls, logits, obj, dep = eng.stage_forward(
            0, feats, mask_preds.float(), proposal_feat.reshape(B, N, C).float(),
            depth_proposal.reshape(B, N, C).float(), H, W)
        return (cls, logits[0], obj.reshape(B, N, C, 1, 1), logits[1], dep.reshape(B, N, C, 1, 1))

    # The reference's KernelUpdateIterHead.get_panoptic calls three helpers on the last stage head
    # (kernel_update.py:423-442 -> kernel_update_head.py:593-650).  Here they are fused into pf_panoptic
    # (postprocess.get_panoptic, called by this package's KernelUpdateIterHead), which never materialises the x4
    # up-sampled maps they return; a caller that mixes the reference's iter head with this stage head gets a clear error.
    def rescale_masks(self, masks_per_img, img_meta):
        _unsupported('KernelUpdateHead.rescale_masks outside polyphonicformer_b200.postprocess.get_panoptic '
                     '(use this package\'s KernelUpdateIterHead, which fuses it into pf_panoptic)')

    def rescale_depth(self, depth, img_meta):
        _unsupported('KernelUpdateHead.rescale_depth outside polyphonicformer_b200.postprocess.get_panoptic '
                     '(use this package\'s KernelUpdateIterHead, which fuses it into pf_panoptic)')

    def segm2result(self, mask_preds, det_labels, cls_scores, depth_preds):
        _unsupported('KernelUpdateHead.segm2result (the non-panoptic result format)')


class KernelUpdateIterHead(nn.Module):
    """Stage loop (reference: polyphonic/kernel_update.py:13-157, 282-401)."""

    def __init__(self, num_stages=6, recursive=False, assign_stages=5, stage_loss_weights=(1, 1, 1, 1, 1, 1),
                 do_panoptic=False, proposal_feature_channel=256, merge_cls_scores=False, post_assign=False,
                 hard_target=False, merge_joint=True, num_proposals=100, num_thing_classes=80, num_stuff_classes=53,
                 mask_assign_stride=4, ignore_label=255, tracking=False, mask_head=None, mask_out_stride=4,
                 train_cfg=None, test_cfg=None, **kwargs):
        super().__init__()
        assert mask_head is not None
        assert len(stage_loss_weights) == num_stages
        self.num_stages = num_stages
        self.stage_loss_weights = stage_loss_weights
        self.proposal_feature_channel = proposal_feature_channel
        self.merge_cls_scores = merge_cls_scores
        self.recursive = recursive
        self.post_assign = post_assign
        self.mask_out_stride = mask_out_stride
        self.hard_target = hard_target
        self.assign_stages = assign_stages
        self.do_panoptic = do_panoptic
        self.merge_joint = merge_joint
        self.num_thing_classes, self.num_stuff_classes = num_thing_classes, num_stuff_classes
        self.mask_assign_stride = mask_assign_stride
        self.num_proposals = num_proposals
        self.ignore_label = ignore_label
        self.tracking = tracking
        self.train_cfg = train_cfg
        self.test_cfg = to_config(test_cfg) if test_cfg is not None else None
        if not isinstance(mask_head, list):
            mask_head = [mask_head for _ in range(num_stages)]
        assert len(mask_head) == num_stages
        self.mask_head = nn.ModuleList([build_head(h) for h in mask_head])
        if recursive:
            for i in range(num_stages):
                self.mask_head[i] = self.mask_head[0]
        self._engine = None

    def init_weights(self):
        for i in range(self.num_stages):
            self.mask_head[i].init_weights()

    @property
    def with_mask(self):
        return True

    def engine(self, device):
        key = (_version_key(self), str(device))
        if self._engine is None or self._engine[0] != key:
            h = self.mask_head[0]
            self._engine = (key, DecoderEngine([m.state_dict() for m in self.mask_head], device, h.num_classes,
                                               h.feedforward_channels))
        return self._engine[1]

    def _mask_forward(self, stage, x, object_feats, mask_preds, img_metas, depth_preds, depth_proposal, depth_feats):
        """kernel_update.py:125-157 -- one stage through KernelUpdateHead.forward (all five outputs)."""
        head = self.mask_head[stage]
        feats = self.mask_head[0]._prepared_feats(x, depth_feats)     # cast once per frame, shared by the stages
        cls_score, mask_preds, object_feats, depth_preds, depth_proposal = head(
            x, object_feats, mask_preds, img_metas=img_metas, depth_preds=depth_preds,
            depth_proposal=depth_proposal, depth_feats=depth_feats, _feats=feats)
        if head.mask_upsample_stride > 1 and (stage == self.num_stages - 1 or self.training):
            if head.mask_upsample_stride != 2:
                _unsupported('mask_upsample_stride=%d' % head.mask_upsample_stride)
            eng = head.engine(x.device)
            scaled_mask_preds = eng.upsample2x(mask_preds)
            scaled_depth_preds = eng.upsample2x(depth_preds)
        else:
            scaled_mask_preds, scaled_depth_preds = mask_preds, depth_preds
        return dict(cls_score=cls_score, mask_preds=mask_preds, scaled_mask_preds=scaled_mask_preds,
                    object_feats=object_feats, scaled_depth_preds=scaled_depth_preds, depth_preds=depth_preds,
                    depth_proposal=depth_proposal)

    def decode(self, x, proposal_feats, mask_preds, depth_feats, depth_proposal, all_stage_outputs=False, upsample=None):
        """The fused stage loop of simple_test (kernel_update.py:316-336): one C call, dead intermediate outputs
        skipped unless ``all_stage_outputs``.  Returns the dict of DecoderEngine.decode (cls_score has the sigmoid)."""
        if not x.is_cuda:
            _unsupported('KernelUpdateIterHead on a %s tensor' % x.device.type)
        head = self.mask_head[-1]
        if head.mask_upsample_stride not in (1, 2):
            _unsupported('mask_upsample_stride=%d' % head.mask_upsample_stride)
        B, N = proposal_feats.shape[:2]
        H, W = x.shape[-2:]
        if mask_preds.shape[-2:] != (H, W):
            _unsupported('mask_preds at a different resolution than the feature map')
        eng = self.engine(x.device)
        feats = self.mask_head[0]._prepared_feats(x, depth_feats)
        out = eng.decode(feats, mask_preds.float(), proposal_feats.reshape(B, N, -1).float(),
                         depth_proposal.reshape(B, N, -1).float(), H, W,
                         upsample=head.mask_upsample_stride == 2 if upsample is None else upsample,
                         all_stage_outputs=all_stage_outputs)
        C = proposal_feats.shape[2]
        out['object_feats'] = out['object_feats'].reshape(B, N, C, 1, 1)
        out['depth_proposal'] = out['depth_proposal'].reshape(B, N, C, 1, 1)
        return out

    def simple_test_mask_preds(self, x, proposal_feats, mask_preds, cls_score, img_metas, depth_preds=None,
                               depth_feats=None, depth_proposal=None, imgs_whwh=None, rescale=False):
        """kernel_update.py:356-401."""
        out = self.decode(x, proposal_feats, mask_preds, depth_feats, depth_proposal)
        return out['object_feats'], out['cls_score'], out['mask_preds'], out['scaled_mask_preds']

    def simple_test(self, x, proposal_feats, mask_preds, cls_score, img_metas, depth_preds=None, depth_feats=None,
                    depth_proposal=None, imgs_whwh=None, aspp_semantic=None, rescale=False, semantic_input=None):
        """kernel_update.py:282-354: stage loop on the GPU kernels, then per-image panoptic merge (postprocess.py)."""
        from . import postprocess
        if not self.do_panoptic:
            raise NotImplementedError
        head = self.mask_head[-1]
        if aspp_semantic is not None:
            raise NotImplementedError('aspp_semantic is not used by the reference either (kernel_update.py:425-426)')
        stride2 = self.mask_head[0].mask_upsample_stride == 2 and head.mask_upsample_stride == 2
        # with the x2 up-sampling on (the shipped configs) the scaled maps are not materialised: pf_panoptic_batch samples
        # the stride-8 logits with the composed taps, bit-identical to upsample -> get_panoptic (kernel_update.py:131-143)
        out = self.decode(x, proposal_feats, mask_preds, depth_feats, depth_proposal, upsample=not stride2)
        depth_initial = depth_preds.detach().float()
        if not stride2 and self.mask_head[0].mask_upsample_stride > 1:
            depth_initial = self.engine(x.device).upsample2x(depth_initial)
        self.last_device_results = []          # per frame: the device tensors behind the returned arrays (tracking path)
        return postprocess.get_panoptic_batch(
            self, head, out['cls_score'], out['mask_preds' if stride2 else 'scaled_mask_preds'], self.test_cfg, img_metas,
            depth_preds=out['depth_preds' if stride2 else 'scaled_depth_preds'], depth_init=depth_initial,
            stride2_inputs=stride2, device_results=self.last_device_results)

    def forward_train(self, *args, **kwargs):
        _unsupported('training (forward_train)')

    def aug_test(self, features, proposal_list, img_metas, rescale=False):
        raise NotImplementedError('SparseMask does not support `aug_test`')


class _ConvGN(nn.Module):
    """Parameters of an mmcv ``ConvModule(C, C, 1, norm_cfg=GN)``: ``conv.weight`` (no bias), ``gn.weight``, ``gn.bias``."""

    def __init__(self, channels, num_groups):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 1, bias=False)
        self.gn = nn.GroupNorm(num_groups, channels)


def _pyramid_supported(fpn):
    """True iff ``fpn`` is a SemanticFPNWrapper in the configuration pf_semantic_fpn implements (the shipped one,
    configs/_base_/models/polyphonic_former.py:78-96): levels 0..3 with 1 / 1 / 2 / 3 [3x3 conv 256->256 without bias + GN32
    + ReLU] modules, the first of level 0 with stride 2, x2 bilinear steps after all but the last conv of levels 2 and 3,
    sine positional encoding (128 feats, normalised) added to level 3, no coordinate channels, sum fusion."""
    try:
        if (fpn.start_level, fpn.end_level, fpn.upsample_times, fpn.cat_coors_level) != (0, 3, 2, 3) or fpn.cat_coors or \
                fpn.fuse_by_cat or len(fpn.convs_all_levels) != 4:
            return False
        pe = fpn.positional_encoding
        if pe is None or (pe.num_feats, bool(pe.normalize), pe.temperature, getattr(pe, 'offset', 0.0)) != (128, True, 10000, 0.0) \
                or abs(pe.scale - 2 * math.pi) > 1e-9 or abs(getattr(pe, 'eps', 1e-6) - 1e-6) > 1e-12:
            return False
        for lvl, n_convs in enumerate((1, 1, 2, 3)):
            seq = fpn.convs_all_levels[lvl]
            names = [n for n, _ in seq.named_children()]
            want = []
            for j in range(n_convs):
                want.append('conv%d' % j)
                if lvl >= 2 and j < n_convs - 1:
                    want.append('upsample%d' % j)
            if names != want:
                return False
            for n, m in seq.named_children():
                if n.startswith('conv'):
                    conv, gn, act = m.conv, m.gn, getattr(m, 'activate', None)
                    stride = (2, 2) if lvl == 0 else (1, 1)
                    if not (tuple(conv.kernel_size) == (3, 3) and tuple(conv.stride) == stride and tuple(conv.padding) == (1, 1)
                            and conv.bias is None and conv.in_channels == conv.out_channels == 256 and conv.groups == 1
                            and tuple(conv.dilation) == (1, 1) and gn.num_groups == 32 and abs(gn.eps - 1e-5) < 1e-12
                            and isinstance(act, nn.ReLU)):
                        return False
                else:
                    if not (m.scale_factor in (2, 2.0) and m.mode == 'bilinear' and m.align_corners is False):
                        return False
        return True
    except AttributeError:
        return False


class _ConvModule(nn.Module):
    """Parameters (and attribute names) of an mmcv ``ConvModule`` with GroupNorm and ReLU: conv (no bias), gn, activate."""

    def __init__(self, k, stride, groups):
        super().__init__()
        self.conv = nn.Conv2d(256, 256, k, stride=stride, padding=k // 2, bias=False)
        self.gn = nn.GroupNorm(groups, 256)
        self.activate = nn.ReLU(inplace=False)


class _SinePE:
    """The attributes of mmdet's SinePositionalEncoding that pf_semantic_fpn bakes in (no parameters)."""

    def __init__(self, num_feats, temperature=10000, normalize=False, scale=2 * math.pi, eps=1e-6, offset=0.0, **kwargs):
        self.num_feats, self.temperature, self.normalize, self.scale, self.eps, self.offset = \
            num_feats, temperature, normalize, scale, eps, offset


class SemanticFPNWrapper(nn.Module):
    """polyphonic/funcs/semantic_fpn.py:16-235 in the shipped configuration, inference only, under the reference's name,
    kwargs and state-dict keys; ``forward`` runs pf_semantic_fpn + pf_fpn_pred.  (Inside ``KernelHead`` the bf16 maps go
    straight on to pf_kernel_head; this ``forward`` returns the reference's list of three fp32 maps.)"""

    def __init__(self, in_channels, feat_channels, out_channels, start_level, end_level, cat_coors=False,
                 positional_encoding=None, cat_coors_level=3, fuse_by_cat=False, return_list=False, upsample_times=3,
                 with_pred=True, num_aux_convs=0, act_cfg=dict(type='ReLU', inplace=True), out_act_cfg=dict(type='ReLU'),
                 conv_cfg=None, norm_cfg=None, **kwargs):
        super().__init__()
        groups = (norm_cfg or {}).get('num_groups')
        if (in_channels, feat_channels, out_channels, start_level, end_level, upsample_times, cat_coors_level, num_aux_convs) != \
                (256, 256, 256, 0, 3, 2, 3, 2) or cat_coors or fuse_by_cat or not with_pred or conv_cfg is not None or \
                (norm_cfg or {}).get('type') != 'GN' or groups != 32 or positional_encoding is None or \
                (act_cfg or {}).get('type') != 'ReLU' or (out_act_cfg or {}).get('type') != 'ReLU':
            _unsupported('SemanticFPNWrapper other than the shipped configuration (configs/_base_/models/polyphonic_former.py:78-96)')
        pe = dict(positional_encoding)
        if pe.pop('type', 'SinePositionalEncoding') != 'SinePositionalEncoding':
            _unsupported('positional_encoding %r' % (positional_encoding,))
        self.in_channels, self.feat_channels, self.out_channels = in_channels, feat_channels, out_channels
        self.start_level, self.end_level, self.upsample_times = start_level, end_level, upsample_times
        self.cat_coors, self.cat_coors_level, self.fuse_by_cat = cat_coors, cat_coors_level, fuse_by_cat
        self.return_list, self.with_pred, self.num_aux_convs = return_list, with_pred, num_aux_convs
        self.positional_encoding = _SinePE(**pe)
        self.convs_all_levels = nn.ModuleList()
        for lvl, n_convs in enumerate((1, 1, 2, 3)):
            seq = nn.Sequential()
            for j in range(n_convs):
                seq.add_module('conv%d' % j, _ConvModule(3, 2 if lvl == 0 else 1, groups))
                if lvl >= 2 and j < n_convs - 1:
                    seq.add_module('upsample%d' % j, nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False))
            self.convs_all_levels.append(seq)
        self.conv_pred = _ConvModule(1, 1, groups)
        self.aux_convs = nn.ModuleList([_ConvModule(1, 1, groups) for _ in range(num_aux_convs)])
        if not _pyramid_supported(self):
            _unsupported('SemanticFPNWrapper: this configuration')
        self._engines = None

    def init_weights(self):
        """semantic_fpn.py:180-185."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, 0, 0.01)

    def forward(self, inputs):
        """semantic_fpn.py:198-235: the four FPN levels -> [conv_pred, aux_convs.0, aux_convs.1] outputs, fp32 [B,256,H,W]."""
        from .kernel_head import FpnPred, SemanticFpnPyramid
        dev = inputs[0].device
        if dev.type != 'cuda':
            _unsupported('SemanticFPNWrapper on a %s device' % dev.type)
        key = (tuple((p.data_ptr(), p._version) for p in self.parameters()), str(dev))
        if self._engines is None or self._engines[0] != key:
            sd = self.state_dict()
            self._engines = (key, SemanticFpnPyramid(sd, dev), FpnPred(sd, dev))
        fused_b, _, hw = self._engines[1].forward(inputs[self.start_level:self.end_level + 1])
        _, maps32 = self._engines[2].forward(fused_b, want_fp32=True, hw=hw)
        return [maps32[0], maps32[1], maps32[2]]


class KernelHead(nn.Module):
    """The proposal stage (reference: polyphonic/kernel_head.py:16-347, 700-706), inference only.

    ``localization_fpn`` (SemanticFPNWrapper) is built through the reference's NECKS registry and runs as the
    reference's PyTorch module (not rebuilt); everything after it -- kernel_head.py:250-336
    -- runs in ``pf_kernel_head`` / ``pf_mask_pool`` / ``pf_init_proposals``.  When the neck exposes the reference's
    ``convs_all_levels`` / ``conv_pred`` / two ``aux_convs`` its last step (semantic_fpn.py:221-229) runs in
    ``pf_fpn_pred`` as well.  ``x_feats`` / ``depth_feats`` are returned as bf16 views of the decoder's feature buffer
    (tagged so that KernelUpdateIterHead uses that buffer directly instead of casting them again)."""

    def __init__(self, num_proposals=100, num_classes=133, num_thing_classes=80, num_stuff_classes=53, in_channels=256,
                 out_channels=256, num_heads=8, num_cls_fcs=1, num_seg_convs=1, num_loc_convs=1, att_dropout=False,
                 localization_fpn=None, conv_kernel_size=1, norm_cfg=dict(type='GN', num_groups=32), semantic_fpn=True,
                 train_cfg=None, xavier_init_kernel=False, kernel_init_std=0.01, use_binary=False,
                 proposal_feats_with_obj=False, loss_mask=None, loss_seg=None, loss_cls=None, loss_dice=None,
                 loss_rank=None, loss_depth=None, feat_downsample_stride=1, feat_refine_stride=1, feat_refine=True,
                 conv_normal_init=False, mask_out_stride=4, hard_target=False, ignore_label=255, cat_stuff_mask=False,
                 with_depth=True, num_depth_convs=1, semantic_out_cfg=None, loss_semantic_seg=None, **kwargs):
        super().__init__()
        self.num_proposals, self.num_classes = num_proposals, num_classes
        self.num_thing_classes, self.num_stuff_classes = num_thing_classes, num_stuff_classes
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_heads, self.num_cls_fcs, self.att_dropout = num_heads, num_cls_fcs, att_dropout
        self.num_loc_convs, self.num_seg_convs, self.num_depth_convs = num_loc_convs, num_seg_convs, num_depth_convs
        self.conv_kernel_size, self.norm_cfg, self.semantic_fpn = conv_kernel_size, norm_cfg, semantic_fpn
        self.train_cfg, self.test_cfg = train_cfg, kwargs.get('test_cfg')
        self.xavier_init_kernel, self.kernel_init_std = xavier_init_kernel, kernel_init_std
        self.use_binary, self.proposal_feats_with_obj = use_binary, proposal_feats_with_obj
        self.feat_downsample_stride, self.feat_refine_stride, self.feat_refine = feat_downsample_stride, feat_refine_stride, feat_refine
        self.conv_normal_init, self.mask_out_stride, self.hard_target = conv_normal_init, mask_out_stride, hard_target
        self.ignore_label, self.cat_stuff_mask, self.with_depth = ignore_label, cat_stuff_mask, with_depth
        self.semantic_out_cfg = semantic_out_cfg
        # reference: x_feats / depth_feats are fp32 NCHW.  Default here: bf16 views of the decoder's feature buffer (the
        # storage precision of the decoder); return_fp32_feats=True returns fp32 tensors instead
        # (the decoder still consumes the shared bf16 buffer through SHARED_FEATS)
        self.return_fp32_feats = bool(kwargs.pop('return_fp32_feats', False))
        for name, cfg in (('loss_mask', loss_mask), ('loss_seg', loss_seg), ('loss_cls', loss_cls), ('loss_dice', loss_dice),
                          ('loss_rank', loss_rank), ('loss_depth', loss_depth), ('loss_semantic_seg', loss_semantic_seg)):
            setattr(self, name, _LossCfg(cfg) if cfg is not None else None)   # inference reads .use_sigmoid only
        if (in_channels, out_channels, conv_kernel_size) != (256, 256, 1):
            _unsupported('KernelHead with in/out channels %d/%d, conv_kernel_size %d' % (in_channels, out_channels, conv_kernel_size))
        if (num_loc_convs, num_seg_convs, num_depth_convs) != (1, 1, 1) or not (semantic_fpn and with_depth):
            _unsupported('KernelHead without exactly one loc / seg / depth conv')
        if feat_downsample_stride > 1 and feat_refine:
            _unsupported('feat_refine (ins_downsample / seg_downsample)')
        if semantic_out_cfg is not None:
            _unsupported('semantic_out_cfg (semantic_aspp)')
        if not (cat_stuff_mask and use_binary and proposal_feats_with_obj):
            _unsupported('KernelHead without cat_stuff_mask / use_binary / proposal_feats_with_obj')
        if norm_cfg.get('type') != 'GN':
            _unsupported('norm_cfg %r' % (norm_cfg,))
        seg_sigmoid = bool(getattr(self.loss_seg, 'use_sigmoid', True)) if self.loss_seg is not None else True
        self.localization_fpn = build_neck(localization_fpn)
        groups = norm_cfg.get('num_groups', 32)
        if groups != 32:
            _unsupported('GroupNorm with num_groups=%d (pf_kernel_head / pf_fpn_pred implement 32 groups of 8 channels)' % groups)
        self.init_kernels = nn.Conv2d(out_channels, num_proposals, 1, bias=False)
        self.conv_seg = nn.Conv2d(out_channels, num_classes if seg_sigmoid else num_classes + 1, 1)
        self.loc_convs = nn.ModuleList([_ConvGN(in_channels, groups)])
        self.seg_convs = nn.ModuleList([_ConvGN(in_channels, groups)])
        self.depth_convs = nn.ModuleList([_ConvGN(in_channels, groups)])
        self.conv_direct_depth = nn.Conv2d(out_channels, 1, 1)
        self._tail = None
        self._fpn_pred = None

    def init_weights(self):
        """kernel_head.py:214-238."""
        if hasattr(self.localization_fpn, 'init_weights'):
            self.localization_fpn.init_weights()
        if self.conv_normal_init:
            for m in (self.loc_convs[0].conv, self.seg_convs[0].conv):
                nn.init.normal_(m.weight, std=0.01)
        prior = -math.log((1 - 0.01) / 0.01)
        nn.init.normal_(self.conv_seg.weight, std=0.01)
        nn.init.constant_(self.conv_seg.bias, prior if self.conv_seg.out_channels == self.num_classes else 0.0)
        if self.xavier_init_kernel:
            nn.init.xavier_uniform_(self.init_kernels.weight)
        else:
            nn.init.normal_(self.init_kernels.weight, mean=0, std=self.kernel_init_std)

    def _own_state(self):
        return {k: v for k, v in self.state_dict().items() if not k.startswith('localization_fpn.')}

    def tail(self, device):
        from .kernel_head import KernelHeadTail
        own = [p for n, p in self.named_parameters() if not n.startswith('localization_fpn.')]
        key = (tuple((p.data_ptr(), p._version) for p in own), str(device))
        if self._tail is None or self._tail[0] != key:
            self._tail = (key, KernelHeadTail(self._own_state(), device, self.num_thing_classes))
        return self._tail[1]

    def _localization_maps(self, img, tail):
        """The three SemanticFPN outputs as bf16 [3][B][256][HWp].  With the reference's SemanticFPNWrapper the pyramid
        (semantic_fpn.py:198-219) stays its PyTorch code and only conv_pred / aux_convs (:221-229) run here."""
        fpn = self.localization_fpn
        def conv_gn_relu_1x1(m):
            """what pf_fpn_pred implements: ConvModule(256, 256, 1, bias=False) + GroupNorm(32, eps=1e-5) + ReLU"""
            conv, gn, act = getattr(m, 'conv', None), getattr(m, 'gn', None), getattr(m, 'activate', None)
            return (isinstance(conv, nn.Conv2d) and tuple(conv.kernel_size) == (1, 1) and conv.bias is None
                    and conv.in_channels == conv.out_channels == 256 and tuple(conv.stride) == (1, 1)
                    and isinstance(gn, nn.GroupNorm) and gn.num_groups == 32 and abs(gn.eps - 1e-5) < 1e-12
                    and isinstance(act, nn.ReLU))

        fused_ok = (all(hasattr(fpn, a) for a in ('convs_all_levels', 'conv_pred', 'aux_convs', 'start_level', 'end_level'))
                    and len(fpn.aux_convs) == 2 and not getattr(fpn, 'fuse_by_cat', False) and getattr(fpn, 'with_pred', True)
                    and all(conv_gn_relu_1x1(m) for m in (fpn.conv_pred, fpn.aux_convs[0], fpn.aux_convs[1])))
        if not fused_ok:
            feats = fpn(img)
            if not isinstance(feats, (list, tuple)) or len(feats) != 3:
                _unsupported('a localization_fpn that does not return [loc, semantic, depth] maps')
            return tail.cast_maps(list(feats)), feats[0].shape[-2:]
        from .kernel_head import FpnPred, SemanticFpnPyramid
        params = list(fpn.parameters())
        key = (tuple((p.data_ptr(), p._version) for p in params), str(img[0].device))
        if self._fpn_pred is None or self._fpn_pred[0] != key:
            sd = fpn.state_dict()
            pred = FpnPred({k: v for k, v in sd.items() if k.startswith('conv_pred.') or k.startswith('aux_convs.')}, img[0].device)
            pyr = SemanticFpnPyramid(sd, img[0].device) if _pyramid_supported(fpn) else None
            self._fpn_pred = (key, pred, pyr)
        _, pred, pyr = self._fpn_pred
        lv = img[fpn.start_level:fpn.end_level + 1]
        if pyr is not None and len(lv) == 4 and lv[1].shape[-2] % 4 == 0 and lv[1].shape[-1] % 4 == 0 and \
                all(tuple(t.shape[-2:]) == (lv[1].shape[-2] * 2 // (1 << i), lv[1].shape[-1] * 2 // (1 << i)) for i, t in enumerate(lv)):
            # the whole neck on the kernels: pf_semantic_fpn (semantic_fpn.py:198-219) -> pf_fpn_pred (:221-229)
            fused_b, _, hw = pyr.forward(lv)
            maps, _ = pred.forward(fused_b, hw=hw)
            return maps, hw
        # a pyramid pf_semantic_fpn does not cover (other depths / strides / coordinate channels, ragged sizes): its 3x3
        # convs stay the module's own PyTorch code, only conv_pred / aux_convs run on the kernels
        if isinstance(fpn, SemanticFPNWrapper):      # this package's neck holds parameters only: there is no PyTorch path
            _unsupported('SemanticFPNWrapper on FPN levels %s (pf_semantic_fpn needs the four levels of a stride-8 map whose '
                         'sides are multiples of 4, which Pad(size_divisor=32) guarantees)' % [tuple(t.shape) for t in lv])
        levels = []
        for i in range(fpn.start_level, fpn.end_level + 1):
            inp = img[i]
            if i == fpn.cat_coors_level:
                if fpn.positional_encoding is not None:
                    inp = inp + fpn.positional_encoding(inp.new_zeros((inp.shape[0],) + inp.shape[-2:], dtype=torch.bool))
                if fpn.cat_coors:
                    inp = torch.cat([inp, fpn.generate_coord(inp)], 1)
            levels.append(fpn.convs_all_levels[i](inp))
        fused = sum(levels)
        maps, _ = pred.forward(fused)
        return maps, fused.shape[-2:]

    def _decode_init_proposals(self, img, img_metas, train_tracking=False):
        """kernel_head.py:240-347 (eval).  Returns the reference's 9-tuple."""
        if self.training:
            _unsupported('KernelHead in training mode')
        dev = next(self.parameters()).device
        if dev.type != 'cuda':
            _unsupported('KernelHead on a %s device' % dev.type)
        tail = self.tail(dev)
        maps, (H, W) = self._localization_maps(img, tail)
        out = tail.forward(maps, H, W, want_fp32_feats=self.return_fp32_feats)
        feats = out['feats']
        B, HW = feats.shape[1], H * W
        if self.return_fp32_feats:
            x_feats, depth_feats = out['x_feats'], out['depth_feats']
        else:
            x_feats = feats[0][..., :HW].reshape(B, 256, H, W)
            depth_feats = feats[1][..., :HW].reshape(B, 256, H, W)
        SHARED_FEATS.publish(x_feats, depth_feats, feats)
        return (out['proposal_feats'], x_feats, out['mask_preds'], None, out['seg_preds'], depth_feats,
                out['depth_proposal'], out['depth_pred'], None)

    def simple_test_rpn(self, img, img_metas, train_tracking=False):
        """kernel_head.py:700-706."""
        return self._decode_init_proposals(img, img_metas, train_tracking)

    def forward_dummy(self, img, img_metas):
        return self._decode_init_proposals(img, img_metas)

    def forward_train(self, *args, **kwargs):
        _unsupported('KernelHead.forward_train (training is out of scope)')
