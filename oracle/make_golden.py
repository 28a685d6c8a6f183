"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the REAL,
unmodified reference (``/root/reference/polyphonic`` + vendored ``mmdet``) on CPU
through ``oracle/mmcv_shim.py``.  Run in the build container only:

    python oracle/make_golden.py

The fixtures hold reference OUTPUTS only; inputs and weights are regenerated
deterministically by ``oracle/synth.py`` on both sides.  ``load_state_dict(strict=True)``
below also proves that the key/shape table in ``oracle/synth.py`` (and hence the
drop-in modules' state-dict layout) equals the reference's.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import mmcv_shim as shim  # noqa: E402
from oracle import synth  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')

DECODER_CASES = [
    # name, B, H, W, seed
    # seed 6: of seeds 0..7 the one whose intermediate mask logits stay furthest from 0 (min |logit| of stages 0/1 =
    # 4.5e-4 / 6.3e-4).  The next stage consumes only sigmoid(logit) > 0.5 (kernel_update_head.py:236-238): a logit
    # within fp32 rounding of 0 (seed 0 has one at -7e-5) flips a mask bit between ANY two implementations, and on a
    # 384-pixel map one flipped pixel moves a pooled row by ~1e-2 -- the un-teacher-forced loop is only comparable
    # away from that discontinuity.  Margins are stored in the fixture ('margin').
    ('decoder_b2_h16_w24_s6', 2, 16, 24, 6),
    ('decoder_b1_h10_w12_s1', 1, 10, 12, 1),   # HW=120: ragged tile tail
]


def build_reference_roi_head():
    shim.install()
    import polyphonic  # noqa: F401  (registers the reference modules)
    from mmdet.models.builder import build_head
    cfg = shim.load_config('/root/reference/configs/polyphonic_image/poly_r50_cityscapes_2x.py')
    roi = cfg.model.roi_head
    roi['train_cfg'] = None
    roi['test_cfg'] = cfg.model.test_cfg.rcnn
    head = build_head(roi)
    head.eval()
    return head


def run_decoder_case(head, B, H, W, seed):
    sd = synth.synth_decoder_state(num_stages=3, seed=seed)
    missing = head.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    inp = synth.synth_decoder_inputs(B, H, W, seed=seed)
    metas = [dict(img_shape=(8 * H, 8 * W, 3), ori_shape=(8 * H, 8 * W, 3), pad_shape=(8 * H, 8 * W, 3),
                  scale_factor=1.0, flip=False, batch_input_shape=(8 * H, 8 * W)) for _ in range(B)]
    out = {}
    with torch.no_grad():
        # kernel_update.py:299-300, 316-328 -- the stage loop of simple_test, verbatim calls
        depth_preds = inp['depth_pred'].expand(-1, inp['depth_proposal'].shape[1], -1, -1)
        object_feats, mask_preds, depth_proposal = inp['proposal_feats'], inp['mask_preds'], inp['depth_proposal']
        for stage in range(head.num_stages):
            r = head._mask_forward(stage, inp['x_feats'], object_feats, mask_preds, metas,
                                   depth_preds, depth_proposal, inp['depth_feats'])
            object_feats, mask_preds = r['object_feats'], r['mask_preds']
            depth_proposal, depth_preds = r['depth_proposal'], r['depth_preds']
            for k in ('cls_score', 'mask_preds', 'object_feats', 'depth_preds', 'depth_proposal'):
                out[f's{stage}.{k}'] = r[k].numpy().astype(np.float32)
        out['scaled_mask_preds'] = r['scaled_mask_preds'].numpy()
        out['scaled_depth_preds'] = r['scaled_depth_preds'].numpy()
        out['cls_score_sigmoid'] = r['cls_score'].sigmoid().numpy()
        # simple_test_mask_preds (kernel_update.py:356-401) must agree with the loop above
        of, cs, mp, smp = head.simple_test_mask_preds(
            inp['x_feats'], inp['proposal_feats'], inp['mask_preds'], None, metas,
            depth_preds=inp['depth_pred'], depth_feats=inp['depth_feats'],
            depth_proposal=inp['depth_proposal'])
        assert torch.equal(smp, r['scaled_mask_preds']) and torch.equal(cs, r['cls_score'].sigmoid())
    return out


def run_updator_case(head, seed=0, rows=37):
    upd = head.mask_head[0].kernel_update_conv
    sd = {k[len('mask_head.0.kernel_update_conv.'):]: v
          for k, v in synth.synth_decoder_state(1, seed).items()
          if k.startswith('mask_head.0.kernel_update_conv.')}
    upd.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(1234 + seed)
    update = torch.randn(rows, synth.C, generator=g) * 20.0
    inputf = torch.randn(rows, 1, synth.C, generator=g)
    with torch.no_grad():
        y = upd(update, inputf)
    return dict(update_feature=update.numpy(), input_feature=inputf.numpy(), out=y.numpy())


PANOPTIC_CASES = [
    # name, h, w (scaled predictions = 1/4 of the padded input), img_shape (crop of the padded input), seed
    ('panoptic_h32_w64_s0', 32, 64, (128, 256), 0),
    ('panoptic_h24_w40_crop_s1', 24, 40, (94, 157), 1),     # KITTI-like: padded 96x160, image 94x157
]


def run_panoptic_case(head, h, w, img_hw, seed):
    """kernel_update.py:421-535 -- the reference's own get_panoptic on hand-constructed predictions."""
    inp = synth.synth_panoptic_inputs(h, w, seed)
    meta = dict(img_shape=(img_hw[0], img_hw[1], 3), ori_shape=(img_hw[0], img_hw[1], 3), pad_shape=(4 * h, 4 * w, 3),
                scale_factor=1.0, flip=False, batch_input_shape=(4 * h, 4 * w))
    with torch.no_grad():
        _, _, (pan, segs), dbasic, dfinal = head.get_panoptic(
            inp['cls_scores'], inp['mask_preds'], head.test_cfg, meta, inp['depth_preds'], inp['depth_init'], None)
    seg = np.array([[s['id'], int(s['isthing']), s['category_id'], s.get('instance_id', -1), s.get('area', -1)]
                    for s in segs], dtype=np.int64)
    score = np.array([s.get('score', -1.0) for s in segs], dtype=np.float64)
    return dict(panoptic=pan, seg=seg, seg_score=score, depth_basic=dbasic, depth_final=dfinal,
                img_hw=np.array(img_hw))


KERNEL_HEAD_CASES = [
    # name, B, H, W: the seed is the one of 0..15 whose initial thing-mask logits stay furthest from 0 (the pooled
    # proposal features only see sigmoid(logit) > 0.5, kernel_head.py:314-317; same reasoning as above)
    ('kernel_head_b2_h16_w24', 2, 16, 24),
    ('kernel_head_b1_h10_w13', 1, 10, 13),   # HW=130: ragged 128-pixel tile, HW % 4 != 0 (no TMA store path)
]


class _FixedFPN(torch.nn.Module):
    def __init__(self, maps):
        super().__init__()
        self.maps = maps

    def forward(self, img):
        return list(self.maps)


def build_reference_rpn_head():
    shim.install()
    import polyphonic  # noqa: F401
    from mmdet.models.builder import build_head
    cfg = shim.load_config('/root/reference/configs/polyphonic_image/poly_r50_cityscapes_2x.py')
    rpn = cfg.model.rpn_head
    rpn['train_cfg'] = None
    rpn['test_cfg'] = cfg.model.test_cfg.rpn
    head = build_head(rpn)
    head.eval()
    return head


def run_kernel_head_case(head, B, H, W):
    """kernel_head.py:240-347 -- the reference's own _decode_init_proposals with SemanticFPN replaced by fixed maps."""
    best = None
    for seed in range(16):
        sd = synth.synth_kernel_head_state(seed)
        r = head.load_state_dict(sd, strict=False)
        assert not r.unexpected_keys and all(k.startswith('localization_fpn.') for k in r.missing_keys), r
        maps = synth.synth_fpn_maps(B, H, W, seed)
        head._modules['localization_fpn'] = _FixedFPN(maps)
        with torch.no_grad():
            out = head._decode_init_proposals(None, [None] * B)
        names = ('proposal_feats', 'x_feats', 'mask_preds', 'cls_scores', 'seg_preds', 'depth_feats', 'depth_proposal',
                 'depth_pred', 'semantic_aspp_out')
        out = {k: v.numpy().astype(np.float32) for k, v in zip(names, out) if v is not None}
        margin = float(np.abs(out['mask_preds'][:, :head.num_proposals]).min())
        if best is None or margin > best[1]:
            best = (seed, margin, out)
    return best


FPN_PRED_CASES = [('fpn_pred_b2_h16_w24_s0', 2, 16, 24, 0), ('fpn_pred_b1_h10_w13_s1', 1, 10, 13, 1)]


def run_fpn_pred_case(rpn, B, H, W, seed):
    """semantic_fpn.py:221-229 -- the reference's own conv_pred / aux_convs modules on a fixed fused map."""
    fpn = rpn.localization_fpn
    sd = synth.synth_fpn_pred_state(seed)
    r = fpn.load_state_dict(sd, strict=False)
    assert not r.unexpected_keys and all(k.startswith('convs_all_levels.') or k.startswith('positional') for k in r.missing_keys), r
    fused = synth.synth_fused_map(B, H, W, seed)
    with torch.no_grad():
        outs = [fpn.conv_pred(fused)] + [conv(fused) for conv in fpn.aux_convs]
    return dict(maps=torch.stack(outs).numpy().astype(np.float32))


SEMANTIC_FPN_CASES = [('semantic_fpn_b1_h16_w24_s0', 1, 16, 24, 0)]


def run_semantic_fpn_case(rpn, B, H, W, seed):
    """semantic_fpn.py:198-235 -- the reference's own SemanticFPNWrapper.forward on synthetic FPN levels."""
    fpn = rpn.localization_fpn
    fpn.load_state_dict(synth.synth_semantic_fpn_state(seed), strict=True)
    with torch.no_grad():
        outs = fpn(synth.synth_fpn_inputs(B, H, W, seed))
    return dict(maps=torch.stack(outs).numpy().astype(np.float32))


def run_tracking_case(seed=0):
    """polyphonic_former_video.py:364-403 + :408-419 on a synthetic clip: the reference's own track_roi_extractor
    (SingleRoIExtractor + RoIAlign), track head, mask -> box helpers and QuasiDenseEmbedTracker, called in the
    reference's order (the detector around them is not needed for this path)."""
    shim.install()
    import polyphonic  # noqa: F401
    from mmdet.models.builder import build_head, build_roi_extractor
    from polyphonic.funcs.utils import tensor_mask2box
    from polyphonic.video.qdtrack.builder import build_tracker
    from polyphonic.video.utils import batch_mask2boxlist, bboxlist2roi
    cfg = shim.load_config('/root/reference/configs/polyphonic_video/poly_r50_cityscapes_1x.py')
    extractor = build_roi_extractor(cfg.model.bbox_roi_extractor)
    head = build_head(cfg.model.track_head)
    head.load_state_dict(synth.synth_track_head_state(seed), strict=True)
    head.eval()
    tracker = build_tracker(cfg.model.tracker)
    out = {}
    with torch.no_grad():
        for t, fr in enumerate(synth.synth_clip(seed=seed)):
            masks = fr['masks'].float()
            rois = bboxlist2roi(batch_mask2boxlist([masks])).clamp(min=0.0)                     # :412-415
            embeds = head(extractor(fr['feats'][:extractor.num_inputs], rois))                  # :416-417
            boxes = torch.zeros((masks.shape[0], 5))
            boxes[:, 4] = fr['scores']
            boxes[:, :4] = torch.tensor(tensor_mask2box(masks))                                 # :386-389
            kept, labels, ids = tracker.match(bboxes=boxes, labels=fr['labels'].long(), track_feats=embeds,
                                              frame_id=t + 1)                                   # :391-396 (cnt starts at 1)
            ids = ids + 1
            ids[ids == -1] = 0                                                                  # :398-399
            out[f'f{t}.rois'] = rois.numpy()
            out[f'f{t}.embeds'] = embeds.numpy()
            out[f'f{t}.boxes'] = kept.numpy()
            out[f'f{t}.labels'] = labels.numpy()
            out[f'f{t}.ids'] = ids.numpy()
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(8)
    head = build_reference_roi_head()
    for name, B, H, W, seed in DECODER_CASES:
        out = run_decoder_case(head, B, H, W, seed)
        path = os.path.join(GOLD, name + '.npz')
        margin = np.array([np.abs(out['s%d.mask_preds' % s]).min() for s in range(3)], dtype=np.float32)
        np.savez_compressed(path, B=B, H=H, W=W, seed=seed, margin=margin, **out)
        print(name, {k: v.shape for k, v in out.items() if k.startswith('s2') or 'scaled' in k},
              f'{os.path.getsize(path) / 1e6:.2f} MB')
    for name, h, w, img_hw, seed in PANOPTIC_CASES:
        out = run_panoptic_case(head, h, w, img_hw, seed)
        np.savez_compressed(os.path.join(GOLD, name + '.npz'), h=h, w=w, seed=seed, **out)
        print(name, out['panoptic'].shape, 'segments', len(out['seg']), np.unique(out['panoptic']))
    rpn = build_reference_rpn_head()
    for name, B, H, W, seed in SEMANTIC_FPN_CASES:
        out = run_semantic_fpn_case(rpn, B, H, W, seed)
        np.savez_compressed(os.path.join(GOLD, name + '.npz'), B=B, H=H, W=W, seed=seed, **out)
        print(name, out['maps'].shape)
    for name, B, H, W, seed in FPN_PRED_CASES:   # before the kernel-head cases replace localization_fpn
        out = run_fpn_pred_case(rpn, B, H, W, seed)
        np.savez_compressed(os.path.join(GOLD, name + '.npz'), B=B, H=H, W=W, seed=seed, **out)
        print(name, out['maps'].shape)
    for name, B, H, W in KERNEL_HEAD_CASES:
        seed, margin, out = run_kernel_head_case(rpn, B, H, W)
        path = os.path.join(GOLD, name + '.npz')
        np.savez_compressed(path, B=B, H=H, W=W, seed=seed, margin=np.float32(margin), **out)
        print(name, 'seed', seed, 'margin', margin, {k: v.shape for k, v in out.items()},
              f'{os.path.getsize(path) / 1e6:.2f} MB')
    trk = run_tracking_case(0)
    np.savez_compressed(os.path.join(GOLD, 'tracking_clip_s0.npz'), seed=0, **trk)
    print('tracking', {k: v.tolist() for k, v in trk.items() if k.endswith('.ids')})
    u = run_updator_case(head)
    np.savez_compressed(os.path.join(GOLD, 'updator_r37_s0.npz'), **u)
    print('updator', u['out'].shape)


if __name__ == '__main__':
    main()
