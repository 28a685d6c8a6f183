"""GPU parity of the tracking path (include/pf_track.h, through the ctypes host module) against the REAL reference's
modules (tests/golden/tracking_clip_s0.npz: SingleRoIExtractor + RoIAlign, QuasiDenseMaskEmbedHeadGTMask, the mask -> box
helpers and QuasiDenseEmbedTracker run by oracle/make_golden.py) and against oracle/tracking_ref.py on random sequences.
Boxes and ids are integer / index work: exact.  Embeddings: 1e-3 relative (north-star gate), measured ~1e-5."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import synth, tracking_ref

pytestmark = pytest.mark.gpu

TRACKER_CFG = dict(init_score_thr=0.35, obj_score_thr=0.3, match_score_thr=0.5, memo_tracklet_frames=5,
                   memo_backdrop_frames=1, memo_momentum=0.8, nms_conf_thr=0.5, nms_backdrop_iou_thr=0.3,
                   nms_class_iou_thr=0.7)     # configs/polyphonic_video/poly_r50_cityscapes_1x.py (oracle defaults)


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    return torch.device('cuda:0')


@pytest.fixture(scope='module')
def engine(dev):
    from polyphonicformer_b200.track import TrackHeadEngine
    return TrackHeadEngine(synth.synth_track_head_state(0), dev)


def test_boxes_match_reference_golden(dev, engine):
    g = np.load(os.path.join(GOLDEN, 'tracking_clip_s0.npz'))
    for t, fr in enumerate(synth.synth_clip(seed=0)):
        rois, tight = engine.boxes_from_masks(fr['masks'].to(dev))
        assert np.allclose(rois.cpu().numpy(), g[f'f{t}.rois'], rtol=0, atol=1e-5), t
        want = torch.stack([tracking_ref.tight_box_of_mask(m) for m in fr['masks'].float()])
        assert torch.equal(tight.cpu(), want), t


def test_boxes_from_panoptic_equal_boxes_from_masks(dev, engine):
    gen = torch.Generator().manual_seed(3)
    H, W = 200, 333
    pan = torch.zeros((H, W), dtype=torch.int32)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing='ij')
    ids = [3, 7, 1, 12, 40]                                   # 40 never occurs: an empty item
    for i, sid in enumerate(ids[:-1]):
        cy, cx = int(torch.randint(20, H - 20, (1,), generator=gen)), int(torch.randint(20, W - 20, (1,), generator=gen))
        pan[((yy - cy).abs() < 15 + 4 * i) & ((xx - cx).abs() < 25 + 7 * i)] = sid
    pan[0, 0] = 9                                              # a segment that is not asked for
    masks = torch.stack([(pan == sid) for sid in ids]).float()
    r1, t1 = engine.boxes_from_masks(masks.to(dev))
    r2, t2 = engine.boxes_from_panoptic(pan.to(dev), ids)
    assert torch.equal(r1, r2) and torch.equal(t1, t2)
    want = torch.stack([tracking_ref.roi_box_of_mask(m) for m in masks]).clamp(min=0)
    assert torch.allclose(r1[:, 1:].cpu(), want, rtol=0, atol=1e-4)
    assert t1[-1].tolist() == [-1.0, -1.0, 10.0, 10.0] and r1[-1].tolist() == [0.0] * 5


def test_roi_features_and_embeddings_match_reference_golden(dev, engine):
    g = np.load(os.path.join(GOLDEN, 'tracking_clip_s0.npz'))
    sd = synth.synth_track_head_state(0)
    for t, fr in enumerate(synth.synth_clip(seed=0)):
        rois = torch.from_numpy(g[f'f{t}.rois']).to(dev)
        feats = [f.to(dev) for f in fr['feats']]
        emb, rf = engine.embed(feats, rois, want_roi_feats=True)
        want_rf = tracking_ref.roi_features(fr['feats'], torch.from_numpy(g[f'f{t}.rois'])[:, 1:])
        l2, mx = rel_err(rf.cpu(), want_rf)
        assert l2 < 1e-5 and mx < 1e-4, ('roi', t, l2, mx)
        l2, mx = rel_err(emb.cpu(), g[f'f{t}.embeds'])
        print('frame %d: embeddings rel err %.2e (l2) %.2e (max)' % (t, l2, mx))
        assert l2 < 1e-3 and mx < 1e-3, ('embed', t, l2, mx)
        emb2 = engine.head(want_rf.to(dev))                    # the head alone on the reference's own RoI features
        l2, mx = rel_err(emb2.cpu(), g[f'f{t}.embeds'])
        assert l2 < 1e-3 and mx < 1e-3, ('head', t, l2, mx)
        # every intermediate layer against the restatement (a wrong GroupNorm would hide behind the FC's averaging)
    x = want_rf
    with torch.no_grad():
        ref_emb = tracking_ref.embed_head(sd, x)
    assert rel_err(engine.head(x.to(dev)).cpu(), ref_emb)[0] < 1e-4


def test_clip_reproduces_reference_track_ids(dev, engine):
    """polyphonic_former_video.py:364-403 for the 4-frame clip: device boxes -> device RoI features + head -> device tracker."""
    from polyphonicformer_b200.track import DeviceTracker
    g = np.load(os.path.join(GOLDEN, 'tracking_clip_s0.npz'))
    tracker = DeviceTracker(dev, **TRACKER_CFG)
    for t, fr in enumerate(synth.synth_clip(seed=0)):
        rois, tight = engine.boxes_from_masks(fr['masks'].to(dev))
        emb = engine.embed([f.to(dev) for f in fr['feats']], rois)
        boxes = torch.cat([tight, fr['scores'].to(dev).view(-1, 1).float()], 1)
        kept, labels, ids = tracker.match(boxes, fr['labels'].to(dev), emb, t + 1)
        ids = ids + 1
        ids[ids == -1] = 0
        assert ids.tolist() == g[f'f{t}.ids'].tolist(), (t, ids.tolist(), g[f'f{t}.ids'].tolist())
        assert np.allclose(kept.cpu().numpy(), g[f'f{t}.boxes'], rtol=0, atol=1e-5)
        assert labels.cpu().tolist() == g[f'f{t}.labels'].tolist()


@pytest.mark.parametrize('seed,backdrop_frames', [(0, 1), (1, 2), (2, 0)])
def test_tracker_random_sequences_match_restatement(dev, seed, backdrop_frames):
    """40 frames of random detections (objects appear, vanish, overlap, change score) through the device tracker and
    through oracle/tracking_ref.QuasiDenseTracker: identical kept order and ids on every frame."""
    from polyphonicformer_b200.track import DeviceTracker
    cfg = dict(TRACKER_CFG, memo_backdrop_frames=backdrop_frames, memo_tracklet_frames=3 + seed)
    ref = tracking_ref.QuasiDenseTracker(**cfg)
    trk = DeviceTracker(dev, **cfg)
    gen = torch.Generator().manual_seed(100 + seed)
    n_obj = 24
    proto = torch.randn(n_obj, 256, generator=gen) * 0.35          # one embedding direction per object
    centre = torch.rand(n_obj, 2, generator=gen) * torch.tensor([600.0, 300.0])
    size = 20 + torch.rand(n_obj, 2, generator=gen) * 60
    label = torch.randint(0, 8, (n_obj,), generator=gen)
    for frame in range(1, 41):
        alive = torch.rand(n_obj, generator=gen) < 0.7
        idx = alive.nonzero().squeeze(1)
        if frame % 13 == 0:
            idx = idx[:0]                                          # a frame without detections
        k = idx.numel()
        c = centre[idx] + frame * 2.0 + torch.randn(k, 2, generator=gen)
        boxes = torch.cat([c - size[idx] / 2, c + size[idx] / 2, torch.rand(k, 1, generator=gen)], 1)
        emb = proto[idx] + 0.05 * torch.randn(k, 256, generator=gen)
        if k == 0:
            continue                                               # the reference does not call match() then (:371, :400)
        want_b, want_l, want_ids = ref.match(boxes, label[idx], emb, frame)
        got_b, got_l, got_ids = trk.match(boxes.to(dev), label[idx].to(dev), emb.to(dev), frame)
        assert got_ids.tolist() == want_ids.tolist(), (frame, got_ids.tolist(), want_ids.tolist())
        assert torch.equal(got_b.cpu(), want_b) and torch.equal(got_l.cpu(), want_l)
    assert ref.next_id > 10


def test_embed_head_at_the_maximum_roi_count_and_on_degenerate_boxes(dev, engine):
    """100 RoIs (max_per_img of the shipped configs: 50 tiles of two RoIs) incl. zero-area boxes at the origin (what an
    empty mask produces), boxes outside the image and boxes of every pyramid level, against the restatement."""
    gen = torch.Generator().manual_seed(9)
    H, W = 256, 512
    feats = [torch.randn(1, 256, H // s, W // s, generator=gen) for s in (4, 8, 16, 32)]
    K = 100
    c = torch.rand(K, 2, generator=gen) * torch.tensor([W * 1.1, H * 1.1])
    half = torch.exp(torch.rand(K, 1, generator=gen) * 6.4) * 0.7 * (0.8 + 0.4 * torch.rand(K, 2, generator=gen))   # 0.6 .. 500 px
    boxes = torch.cat([c - half, c + half], 1).clamp(min=0)
    boxes[7] = 0.0                                                                   # an empty mask's RoI
    boxes[8] = torch.tensor([W + 50.0, H + 50.0, W + 90.0, H + 70.0])               # entirely outside
    rois = torch.cat([torch.zeros(K, 1), boxes], 1)
    emb, rf = engine.embed([f.to(dev) for f in feats], rois.to(dev), want_roi_feats=True)
    want_rf = tracking_ref.roi_features(feats, boxes)
    l2, mx = rel_err(rf.cpu(), want_rf)
    assert l2 < 1e-5 and mx < 1e-4, (l2, mx)
    with torch.no_grad():
        want = tracking_ref.embed_head(synth.synth_track_head_state(0), want_rf)
    l2, mx = rel_err(emb.cpu(), want)
    print('100 RoIs: embeddings rel err %.2e (l2) %.2e (max)' % (l2, mx))
    assert l2 < 1e-3 and mx < 1e-3, (l2, mx)
    lv = torch.floor(torch.log2(torch.sqrt((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])) / 56 + 1e-6)).clamp(0, 3)
    assert set(lv.long().tolist()) == {0, 1, 2, 3}


def test_tracker_memo_overflow_is_reported_not_silent(dev):
    """More live tracklets than PF_TRACK_MAX_TRACKS (512): the kernel drops the surplus and raises the flag; the host wrapper
    turns it into an error instead of returning wrong ids."""
    from polyphonicformer_b200 import _cabi
    from polyphonicformer_b200.track import DeviceTracker
    trk = DeviceTracker(dev, **dict(TRACKER_CFG, memo_tracklet_frames=100))
    gen = torch.Generator().manual_seed(2)
    with pytest.raises(_cabi.PFError):
        for frame in range(1, 8):                       # 7 x 100 new, mutually distinct tracks
            boxes = torch.cat([torch.arange(100.0).view(-1, 1) * 30 + torch.tensor([[0.0, 0.0, 20.0, 20.0]]),
                               0.9 - torch.arange(100.0).view(-1, 1) * 1e-3], 1)
            emb = torch.randn(100, 256, generator=gen) * 3
            _, _, ids = trk.match(boxes.to(dev), torch.zeros(100, dtype=torch.long, device=dev) + frame, emb.to(dev), frame)
            assert ids.min() >= 0 and len(set(ids.tolist())) == 100
