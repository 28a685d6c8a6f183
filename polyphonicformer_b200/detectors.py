"""The two detectors of the reference under their own names, inference only:

  Polyphonic        polyphonic/polyphonic_former.py:10-161   (image model)
  PolyphonicVideo   polyphonic/polyphonic_former_video.py:16-451 (video model: the image model + tracking head)

plus the drop-in modules of the tracking path the video config names (configs/polyphonic_video/poly_r50_cityscapes_1x.py):
``QuasiDenseMaskEmbedHeadGTMask`` (track_head), ``QuasiDenseEmbedTracker`` (tracker) and ``SingleRoIExtractor``
(bbox_roi_extractor), all on the kernels of include/pf_track.h.

These are thin orchestration, as the reference's are (mmdet ``TwoStageDetector`` plumbing, two_stage.py:37-50): the
backbone and the FPN neck stay the reference's PyTorch modules (north-star) and are built through mmdet's registries when a
reference checkout is importable, or passed in as already-built ``nn.Module``s; the heads are this package's.  There is
no PyTorch fallback for any head: they raise without an sm_100 device.
"""
import torch
import torch.nn as nn

from . import _cabi
from .registry import ConfigDict, MODELS, TRACKERS, build_head, to_config
from .track import DeviceTracker, TrackHeadEngine, paint_maps


def _unsupported(what):
    raise NotImplementedError('%s is outside the B200 inference path (SURVEY.md section 8: training, losses and '
                              'augmentation stay with the reference)' % what)


def _build_external(cfg, kind):
    """backbone / neck: an already-built module, or a config for mmdet's registry (reference checkout importable)."""
    if cfg is None or isinstance(cfg, nn.Module):
        return cfg
    try:
        from mmdet.models import builder as mm
    except ImportError:
        return MODELS.build(cfg)          # someone registered the type locally; KeyError names it otherwise
    return getattr(mm, 'build_' + kind)(cfg)


# ----------------------------------------------------------------------------------------------- tracking drop-ins
class _ConvGN3(nn.Module):
    """Parameters of an mmcv ``ConvModule(256, 256, 3, padding=1, norm_cfg=GN32)``: conv.weight (no bias), gn.weight, gn.bias."""

    def __init__(self, cin, cout, num_groups):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 3, padding=1, bias=False)
        self.gn = nn.GroupNorm(num_groups, cout)


class QuasiDenseMaskEmbedHeadGTMask(nn.Module):
    """polyphonic/video/track_heads.py:13-102, inference: 4 x [3x3 conv + GN + ReLU] on 7x7 RoI features, FC + ReLU, FC."""

    def __init__(self, num_convs=4, num_fcs=1, roi_feat_size=7, in_channels=256, conv_out_channels=256,
                 fc_out_channels=1024, embed_channels=256, conv_cfg=None, norm_cfg=None, softmax_temp=-1, loss_track=None,
                 loss_track_aux=None, **kwargs):
        super().__init__()
        ng = (norm_cfg or {}).get('num_groups')
        if (num_convs, num_fcs, roi_feat_size, in_channels, conv_out_channels, fc_out_channels, embed_channels) != \
                (4, 1, 7, 256, 256, 1024, 256) or conv_cfg is not None or (norm_cfg or {}).get('type') != 'GN' or ng != 32:
            _unsupported('QuasiDenseMaskEmbedHeadGTMask other than the shipped configuration (4 convs + GN32, 1 FC of 1024, '
                         '7x7 RoIs, 256 channels)')
        self.num_convs, self.num_fcs, self.roi_feat_size = num_convs, num_fcs, roi_feat_size
        self.in_channels, self.conv_out_channels = in_channels, conv_out_channels
        self.fc_out_channels, self.embed_channels = fc_out_channels, embed_channels
        self.softmax_temp = softmax_temp
        self.convs = nn.ModuleList([_ConvGN3(in_channels if i == 0 else conv_out_channels, conv_out_channels, ng)
                                    for i in range(num_convs)])
        self.fcs = nn.ModuleList([nn.Linear(conv_out_channels * roi_feat_size ** 2, fc_out_channels)])
        self.fc_embed = nn.Linear(fc_out_channels, embed_channels)
        self._engine = None

    def init_weights(self):
        for m in self.fcs:
            nn.init.xavier_uniform_(m.weight)
            nn.init.constant_(m.bias, 0)
        nn.init.normal_(self.fc_embed.weight, 0, 0.01)
        nn.init.constant_(self.fc_embed.bias, 0)

    def engine(self, device, strides=(4, 8, 16, 32)):
        key = (tuple((p.data_ptr(), p._version) for p in self.parameters()), str(device), tuple(strides))
        if self._engine is None or self._engine[0] != key:
            self._engine = (key, TrackHeadEngine(self.state_dict(), device, strides))
        return self._engine[1]

    def forward(self, x):
        """RoI features [K,256,7,7] -> embeddings [K,256] (track_heads.py:92-102)."""
        if self.training:
            _unsupported('QuasiDenseMaskEmbedHeadGTMask in training mode')
        if not x.is_cuda:
            raise _cabi.PFError(-4, 'QuasiDenseMaskEmbedHeadGTMask', 'needs a CUDA (sm_100) device; there is no CPU fallback')
        if x.shape[0] == 0:
            return x.new_zeros((0, self.embed_channels))
        return self.engine(x.device).head(x)


class SingleRoIExtractor(nn.Module):
    """The ``bbox_roi_extractor`` of the video config (mmdet SingleRoIExtractor + mmcv RoIAlign).  Here it only carries
    the geometry: PolyphonicVideo runs RoIAlign fused with the embedding head (pf_track_embed)."""

    def __init__(self, roi_layer, out_channels, featmap_strides, finest_scale=56, **kwargs):
        super().__init__()
        rl = dict(roi_layer)
        if rl.get('type') != 'RoIAlign' or rl.get('output_size') != 7 or rl.get('sampling_ratio', 0) != 2 or \
                out_channels != 256 or finest_scale != 56 or len(featmap_strides) != 4:
            _unsupported('a RoI extractor other than RoIAlign 7x7, sampling_ratio 2, 4 levels, finest_scale 56')
        self.featmap_strides = list(featmap_strides)
        self.out_channels = out_channels

    @property
    def num_inputs(self):
        return len(self.featmap_strides)


class QuasiDenseEmbedTracker(object):
    """polyphonic/video/qdtrack/trackers/quasi_dense_embed_tracker.py with the memo on the device (pf_tracker_match).
    Built from the config without a device, like the reference's; the device is the one of the first ``match`` call."""

    def __init__(self, **cfg):
        self.cfg = dict(cfg)
        if self.cfg.get('match_metric', 'bisoftmax') != 'bisoftmax':
            _unsupported('match_metric=%r' % self.cfg['match_metric'])
        self._dev = None

    def match(self, bboxes, labels, track_feats, frame_id, asso_tau=-1):
        if self._dev is None or self._dev.device != bboxes.device:
            self._dev = DeviceTracker(bboxes.device, **self.cfg)
        return self._dev.match(bboxes, labels, track_feats, frame_id)

    @property
    def device_tracker(self):
        return self._dev


# ----------------------------------------------------------------------------------------------- detectors
class Polyphonic(nn.Module):
    """polyphonic/polyphonic_former.py:10-161 on mmdet's TwoStageDetector plumbing (two_stage.py:22-60), inference."""

    def __init__(self, backbone, neck=None, rpn_head=None, roi_head=None, train_cfg=None, test_cfg=None, pretrained=None,
                 init_cfg=None, num_thing_classes=80, num_stuff_classes=53, mask_assign_stride=4, semantic_kitti=False):
        super().__init__()
        self.backbone = _build_external(backbone, 'backbone')
        self.neck = _build_external(neck, 'neck')
        train_cfg = to_config(train_cfg) if train_cfg is not None else None
        test_cfg = to_config(test_cfg) if test_cfg is not None else ConfigDict(rpn=None, rcnn=None)
        assert rpn_head is not None, 'KNet does not support external proposals'
        if isinstance(rpn_head, nn.Module):
            self.rpn_head = rpn_head
        else:
            cfg = dict(rpn_head)
            cfg.update(train_cfg=train_cfg.rpn if train_cfg is not None else None, test_cfg=test_cfg.rpn)   # two_stage.py:37-41
            self.rpn_head = build_head(cfg)
        if isinstance(roi_head, nn.Module):
            self.roi_head = roi_head
        else:
            cfg = dict(roi_head)
            cfg.update(train_cfg=train_cfg.rcnn if train_cfg is not None else None, test_cfg=test_cfg.rcnn)  # :43-50
            self.roi_head = build_head(cfg)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.num_thing_classes, self.num_stuff_classes = num_thing_classes, num_stuff_classes
        self.mask_assign_stride, self.semantic_kitti = mask_assign_stride, semantic_kitti

    with_rpn = with_roi_head = True

    @property
    def with_neck(self):
        return self.neck is not None

    def extract_feat(self, img):
        """two_stage.py:78-84: backbone + neck (the reference's PyTorch modules)."""
        x = self.backbone(img)
        if self.with_neck:
            x = self.neck(x)
        return x

    def forward_train(self, *args, **kwargs):
        _unsupported('Polyphonic.forward_train')

    def _decode(self, img, img_metas, rescale=False):
        x = self.extract_feat(img)
        rpn_results = self.rpn_head.simple_test_rpn(x, img_metas)
        (proposal_feats, x_feats, mask_preds, cls_scores, seg_preds, depth_feats, depth_proposal, depth_pred,
         semantic_aspp_out) = rpn_results
        segm_results = self.roi_head.simple_test(
            x_feats, proposal_feats, mask_preds, cls_scores, img_metas, depth_preds=depth_pred, depth_feats=depth_feats,
            depth_proposal=depth_proposal, imgs_whwh=None, aspp_semantic=semantic_aspp_out, rescale=rescale)
        return x, segm_results

    def simple_test(self, img, img_metas, proposals=None, rescale=False):
        """polyphonic_former.py:130-161."""
        return self._decode(img, img_metas, rescale)[1]

    def forward_test(self, imgs, img_metas, **kwargs):
        """mmdet BaseDetector.forward_test (base.py:113-149) without test-time augmentation."""
        if len(imgs) != 1 or len(img_metas) != 1:
            _unsupported('test-time augmentation (aug_test)')
        for img, metas in zip(imgs, img_metas):
            for m in metas:
                m['batch_input_shape'] = tuple(img.size()[-2:])
        return self.simple_test(imgs[0], img_metas[0], **kwargs)

    def forward(self, img, img_metas, return_loss=True, **kwargs):
        if return_loss:
            return self.forward_train(img, img_metas, **kwargs)
        return self.forward_test(img, img_metas, **kwargs)


class PolyphonicVideo(Polyphonic):
    """polyphonic/polyphonic_former_video.py:16-451, inference: the image model, then per frame thing masks -> boxes ->
    RoI features -> embeddings -> association with the tracker memo -> track-id / semantic maps."""

    def __init__(self, *args, track_head=None, bbox_roi_extractor=None, track_train_cfg=None, tracker=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.num_proposals = self.rpn_head.num_proposals
        self.tracker = None
        self.cnt = -1
        if track_head is not None:
            self.track_train_cfg = track_train_cfg          # assigner / sampler: training only, not built
            self.track_head = track_head if isinstance(track_head, nn.Module) else build_head(track_head)
            self.track_roi_extractor = (bbox_roi_extractor if isinstance(bbox_roi_extractor, nn.Module)
                                        else MODELS.build(bbox_roi_extractor))
        if tracker is not None:
            self.tracker_cfg = tracker

    def init_tracker(self):
        """polyphonic_former_video.py:58-60."""
        self.tracker = TRACKERS.build(self.tracker_cfg) if isinstance(self.tracker_cfg, dict) else self.tracker_cfg
        self.cnt = 1

    # ---- per-frame pieces (also used frame-sharded by video.VideoPipeline)
    @staticmethod
    def get_things_id_for_tracking(segments_info):
        """:421-434 without the K full-resolution masks: they are `panoptic == id`, which the kernels test directly."""
        things = [s for s in segments_info if s['isthing']]
        return ([s['instance_id'] for s in things], [s['category_id'] for s in things], [s['id'] for s in things],
                [s['score'] for s in things])

    def _track_forward(self, x, mask_pred):
        """:408-419 for the masks [K,H,W] of one frame (test mode)."""
        eng = self.track_head.engine(mask_pred.device, self.track_roi_extractor.featmap_strides)
        rois, _ = eng.boxes_from_masks(mask_pred)
        return eng.embed(x[:self.track_roi_extractor.num_inputs], rois)

    def track_records(self, x, panoptic_dev, segments_info):
        """Everything of :364-390 that does not need the memo: (seg_ids, bboxes [K,5], labels [K], embeds [K,256]) on the
        device, or None without thing segments."""
        _, labels, seg_ids, scores = self.get_things_id_for_tracking(segments_info)
        if not labels:
            return None
        dev = panoptic_dev.device
        eng = self.track_head.engine(dev, self.track_roi_extractor.featmap_strides)
        rois, tight = eng.boxes_from_panoptic(panoptic_dev, seg_ids)
        embeds = eng.embed(x[:self.track_roi_extractor.num_inputs], rois)
        bboxes = torch.cat([tight, torch.tensor(scores, dtype=torch.float32, device=dev).view(-1, 1)], 1)
        return seg_ids, bboxes, torch.tensor(labels, dtype=torch.int64, device=dev), embeds

    def paint(self, panoptic_dev, segments_info, seg_ids, ids):
        """generate_track_id_maps + get_semantic_seg (:436-451), see track.paint_maps."""
        return paint_maps(panoptic_dev, segments_info, seg_ids, ids, self.num_thing_classes + self.num_stuff_classes)

    def simple_test(self, img, img_metas, proposals=None, rescale=False):
        """polyphonic_former_video.py:326-403 (bs = 1, as the reference: `results = segm_results[0]`)."""
        x, segm_results = self._decode(img, img_metas, rescale)
        _, _, (panoptic_seg, segments_info), _, depth_final = segm_results[0]
        dev_res = getattr(self.roi_head, 'last_device_results', None)
        pan_dev = dev_res[0]['panoptic'] if dev_res else torch.from_numpy(panoptic_seg).to(img.device)
        rec = self.track_records(x, pan_dev, segments_info)
        ids, seg_ids = [], []
        if rec is not None:
            seg_ids, bboxes, labels, embeds = rec
            assert self.cnt > 0, 'init_tracker() has not been called'
            _, _, ids = self.tracker.match(bboxes=bboxes, labels=labels, track_feats=embeds, frame_id=self.cnt)
            self.cnt += 1
            ids = ids + 1
            ids[ids == -1] = 0
            ids = ids.tolist()
        sem, trk = self.paint(pan_dev, segments_info, seg_ids, ids)
        return [{'sem': sem, 'track': trk, 'depth': depth_final}]
