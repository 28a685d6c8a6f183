"""The video model's per-frame glue on the device (polyphonic/polyphonic_former_video.py:326-451) against
oracle/tracking_ref.py: PolyphonicVideo.simple_test frame by frame, and the frame-sharded VideoShardRunner (one rank here;
scripts/video_shard_check.py runs the same check under torchrun with NCCL)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import synth, tracking_ref

pytestmark = pytest.mark.gpu
NUM_THING, NUM_STUFF = 8, 11


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    return torch.device('cuda:0')


def video_cfg():
    return json.load(open(os.path.join(GOLDEN, 'video_cfg.json')))


def clip_frames(seed=0, n_frames=4):
    """synth_clip with its (overlapping) masks turned into a panoptic map + segments_info the way get_panoptic would hand
    them over: segment ids 1.. in descending score order, later segments never overwrite earlier ones, plus one stuff
    segment on the remaining pixels of the top rows."""
    out = []
    for fr in synth.synth_clip(n_frames=n_frames, seed=seed):
        order = torch.argsort(fr['scores'], descending=True).tolist()
        H, W = fr['masks'].shape[1:]
        pan = torch.zeros((H, W), dtype=torch.int32)
        info = []
        for k in order:
            free = fr['masks'][k] & (pan == 0)
            if free.sum() == 0:
                continue
            sid = len(info) + 1
            pan[free] = sid
            info.append({'id': sid, 'isthing': True, 'score': float(fr['scores'][k]), 'category_id': int(fr['labels'][k]),
                         'instance_id': sid})
        sid = len(info) + 1
        stuff = pan == 0
        stuff[10:] = False
        pan[stuff] = sid
        info.append({'id': sid, 'isthing': False, 'category_id': NUM_THING + 3, 'area': int(stuff.sum())})
        out.append(dict(feats=fr['feats'], panoptic=pan, info=info, depth=torch.rand(H, W).numpy()))
    return out


def expected_maps(tracker, sd, fr, frame_id):
    """The reference's glue (:364-403, :436-451) on the oracle's pieces."""
    things = [s for s in fr['info'] if s['isthing']]
    masks = torch.stack([fr['panoptic'] == s['id'] for s in things])
    labels = torch.tensor([s['category_id'] for s in things])
    scores = torch.tensor([s['score'] for s in things])
    ids, _ = tracking_ref.track_frame(sd, tracker, fr['feats'], masks, labels, scores, frame_id)
    track = np.zeros(fr['panoptic'].shape)
    for i, tid in enumerate(ids.tolist()):                    # ids[i] onto masks[i]: kept order onto segment order
        track[masks[i].numpy()] = tid
    sem = np.full(fr['panoptic'].shape, NUM_THING + NUM_STUFF, dtype=np.uint8)
    for s in fr['info']:
        sem[(fr['panoptic'] == s['id']).numpy()] = s['category_id']
    return ids.tolist(), track, sem


class _Backbone(torch.nn.Module):
    """Stands in for ResNet + FPN: hands out the prepared pyramid of the current frame."""

    def forward(self, img):
        return self.levels


class _Rpn(torch.nn.Module):
    num_proposals = 100

    def simple_test_rpn(self, x, img_metas):
        return (None,) * 9


class _Roi(torch.nn.Module):
    """Stands in for KernelUpdateIterHead.simple_test: hands out the prepared panoptic result (host + device copies)."""

    def simple_test(self, *args, **kwargs):
        fr = self.frame
        self.last_device_results = [dict(panoptic=fr['panoptic'].to(self.dev))]
        return [(None, None, (fr['panoptic'].numpy(), fr['info']), None, fr['depth'])]


def test_video_model_simple_test_matches_restatement(dev):
    from polyphonicformer_b200 import registry
    cfg = video_cfg()
    backbone, roi = _Backbone(), _Roi()
    roi.dev = dev
    model = registry.build_detector(dict(type='PolyphonicVideo', backbone=backbone, neck=None, rpn_head=_Rpn(), roi_head=roi,
                                         num_thing_classes=NUM_THING, num_stuff_classes=NUM_STUFF, track_head=cfg['track_head'],
                                         bbox_roi_extractor=cfg['bbox_roi_extractor'], tracker=cfg['tracker'],
                                         track_train_cfg=cfg['track_train_cfg']))
    sd = synth.synth_track_head_state(0)
    model.track_head.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    ref = tracking_ref.QuasiDenseTracker(**{k: v for k, v in cfg['tracker'].items() if k not in ('type', 'with_cats', 'match_metric')})
    seen = set()
    for clip in range(2):                                       # the second clip restarts the tracker (video_inference.py:24-25)
        model.init_tracker()
        ref = tracking_ref.QuasiDenseTracker(**{k: v for k, v in cfg['tracker'].items()
                                                if k not in ('type', 'with_cats', 'match_metric')})
        for t, fr in enumerate(clip_frames(seed=clip)):
            backbone.levels = [f.to(dev) for f in fr['feats']]
            roi.frame = fr
            out = model.simple_test(torch.zeros(1, 3, 8, 8, device=dev), [{}])
            ids, track, sem = expected_maps(ref, sd, fr, t + 1)
            assert len(out) == 1 and set(out[0]) == {'sem', 'track', 'depth'}
            assert out[0]['track'].dtype == np.float64 and np.array_equal(out[0]['track'], track), (clip, t)
            assert out[0]['sem'].dtype == np.uint8 and np.array_equal(out[0]['sem'], sem)
            assert out[0]['depth'] is fr['depth']
            seen.update(ids)
        assert model.cnt == 5
    assert len(seen) >= 5 and 0 in seen


def test_shard_runner_single_rank_matches_restatement(dev):
    """VideoShardRunner's tracking half (records -> gather -> replay -> paint) over two 4-frame clips in waves of 3
    frames, so that waves straddle the clip boundary."""
    from types import SimpleNamespace
    from polyphonicformer_b200.track import TrackHeadEngine
    from polyphonicformer_b200.video import VideoShardRunner
    cfg = video_cfg()
    sd = synth.synth_track_head_state(0)
    runner = VideoShardRunner(SimpleNamespace(device=dev), TrackHeadEngine(sd, dev), None, None, None, cfg['tracker'], NUM_THING,
                              NUM_STUFF, clip_len=4)
    frames = clip_frames(seed=0) + clip_frames(seed=1)
    frames.append(dict(frames[0], info=[s for s in frames[0]['info'] if not s['isthing']]))     # a 9th frame without things
    # ... and frames with exactly ONE thing: a one-row record is a "contiguous" view at an unaligned offset of the gathered
    # buffer (this crashed the 8-rank bench once)
    one = [s for s in frames[1]['info'] if s['isthing']][:1] + [s for s in frames[1]['info'] if not s['isthing']]
    frames += [dict(frames[1], info=one), dict(frames[2], info=one), dict(frames[3], info=one)]
    tcfg = {k: v for k, v in cfg['tracker'].items() if k not in ('type', 'with_cats', 'match_metric')}
    ref = None
    want = []
    cnt = 1
    for g, fr in enumerate(frames):
        if g % 4 == 0:
            ref, cnt = tracking_ref.QuasiDenseTracker(**tcfg), 1
        if any(s['isthing'] for s in fr['info']):
            want.append(expected_maps(ref, sd, fr, cnt))
            cnt += 1
        else:
            sem = np.full(fr['panoptic'].shape, NUM_THING + NUM_STUFF, dtype=np.uint8)
            for s in fr['info']:
                sem[(fr['panoptic'] == s['id']).numpy()] = s['category_id']
            want.append(([], np.zeros(fr['panoptic'].shape), sem))
    for wave in range(4):
        gids = runner.global_ids(wave, 3)
        assert gids == [3 * wave, 3 * wave + 1, 3 * wave + 2]
        local = [frames[g] for g in gids]
        res = [(None, None, (fr['panoptic'].numpy(), fr['info']), None, fr['depth']) for fr in local]
        dev_res = [dict(panoptic=fr['panoptic'].to(dev)) for fr in local]
        fpn = [torch.stack([fr['feats'][l][0] for fr in local]).to(dev) for l in range(4)]
        recs = runner.track_records(res, dev_res, fpn)
        ids = runner.associate(recs, gids)
        out = runner.paint(dev_res, res, recs, gids, ids)
        for g, o in zip(gids, out):
            assert ids.get(g, []) == want[g][0], (g, ids.get(g), want[g][0])
            assert np.array_equal(o['track'], want[g][1]) and np.array_equal(o['sem'], want[g][2]), g
