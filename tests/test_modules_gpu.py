"""GPU parity of the drop-in modules (reference-facing API) against the golden outputs of the real reference."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_err
from oracle import synth

pytestmark = pytest.mark.gpu
GATE, TIGHT = 1e-3, 5e-5


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    return torch.device('cuda:0')


def build_roi_head(dev, seed):
    import polyphonicformer_b200 as pf
    d = json.load(open(os.path.join(GOLDEN, 'roi_head_cfg.json')))
    head = pf.build_head(dict(d['roi_head'], train_cfg=None, test_cfg=d['test_cfg']))
    head.load_state_dict(synth.synth_decoder_state(3, seed), strict=True)
    return head.to(dev).eval()


def metas(B, H, W):
    return [dict(img_shape=(8 * H, 8 * W, 3), ori_shape=(8 * H, 8 * W, 3), pad_shape=(8 * H, 8 * W, 3),
                 scale_factor=1.0, flip=False, batch_input_shape=(8 * H, 8 * W)) for _ in range(B)]


def test_kernel_updator_module_matches_reference_golden(dev):
    import polyphonicformer_b200 as pf
    g = np.load(os.path.join(GOLDEN, 'updator_r37_s0.npz'))
    upd = pf.build_transformer_layer(dict(type='KernelUpdator', in_channels=256, feat_channels=256, out_channels=256,
                                          input_feat_shape=3, act_cfg=dict(type='ReLU', inplace=True),
                                          norm_cfg=dict(type='LN')))
    sd = {k[len('mask_head.0.kernel_update_conv.'):]: v for k, v in synth.synth_decoder_state(1, 0).items()
          if k.startswith('mask_head.0.kernel_update_conv.')}
    upd.load_state_dict(sd, strict=True)
    upd = upd.to(dev)
    with torch.no_grad():
        y = upd(torch.from_numpy(g['update_feature']).to(dev), torch.from_numpy(g['input_feature']).to(dev))
    torch.cuda.synchronize()
    assert y.shape == (37, 1, 256)
    l2, mx = rel_err(y.cpu(), g['out'])
    assert l2 < TIGHT and mx < TIGHT, (l2, mx)


@pytest.mark.parametrize('name', ['decoder_b2_h16_w24_s6', 'decoder_b1_h10_w12_s1'])
def test_mask_forward_per_stage_matches_reference_golden(dev, name):
    """KernelUpdateIterHead._mask_forward with fp32 NCHW inputs exactly as the reference calls it."""
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    B, H, W, seed = int(g['B']), int(g['H']), int(g['W']), int(g['seed'])
    head = build_roi_head(dev, seed)
    inp = {k: v.to(dev) for k, v in synth.synth_decoder_inputs(B, H, W, seed).items()}
    obj, mask, dprop = inp['proposal_feats'], inp['mask_preds'], inp['depth_proposal']
    dpred = inp['depth_pred'].expand(-1, dprop.shape[1], -1, -1)
    with torch.no_grad():
        for s in range(3):
            r = head._mask_forward(s, inp['x_feats'], obj, mask, metas(B, H, W), dpred, dprop, inp['depth_feats'])
            torch.cuda.synchronize()
            for k in ('cls_score', 'mask_preds', 'object_feats', 'depth_preds', 'depth_proposal'):
                assert r[k].shape == g['s%d.%s' % (s, k)].shape
                l2, mx = rel_err(r[k].cpu(), g['s%d.%s' % (s, k)])
                assert l2 < TIGHT and mx < TIGHT, (name, s, k, l2, mx)
            # teacher forcing with the reference's outputs
            obj = torch.from_numpy(g['s%d.object_feats' % s]).to(dev)
            mask = torch.from_numpy(g['s%d.mask_preds' % s]).to(dev)
            dprop = torch.from_numpy(g['s%d.depth_proposal' % s]).to(dev)
            dpred = torch.from_numpy(g['s%d.depth_preds' % s]).to(dev)
        for k in ('scaled_mask_preds', 'scaled_depth_preds'):
            l2, mx = rel_err(r[k].cpu(), g[k])
            assert l2 < TIGHT and mx < TIGHT, (name, k, l2, mx)


def test_simple_test_mask_preds_and_simple_test(dev):
    name = 'decoder_b2_h16_w24_s6'
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    B, H, W, seed = int(g['B']), int(g['H']), int(g['W']), int(g['seed'])
    head = build_roi_head(dev, seed)
    inp = {k: v.to(dev) for k, v in synth.synth_decoder_inputs(B, H, W, seed).items()}
    with torch.no_grad():
        of, cs, mp, smp = head.simple_test_mask_preds(
            inp['x_feats'], inp['proposal_feats'], inp['mask_preds'], None, metas(B, H, W),
            depth_preds=inp['depth_pred'], depth_feats=inp['depth_feats'], depth_proposal=inp['depth_proposal'])
        torch.cuda.synchronize()
        for got, key in ((of, 's2.object_feats'), (cs, 'cls_score_sigmoid'), (mp, 's2.mask_preds'),
                         (smp, 'scaled_mask_preds')):
            assert got.shape == g[key].shape
            l2, mx = rel_err(got.cpu(), g[key])
            assert l2 < GATE and mx < GATE, (key, l2, mx)
        res = head.simple_test(inp['x_feats'], inp['proposal_feats'], inp['mask_preds'], None, metas(B, H, W),
                               depth_preds=inp['depth_pred'], depth_feats=inp['depth_feats'],
                               depth_proposal=inp['depth_proposal'])
    assert len(res) == B
    for r in res:
        assert r[0] is None and r[1] is None
        pan, segs = r[2]
        assert pan.shape == (8 * H, 8 * W) and pan.dtype == np.int32
        assert r[3].shape == (8 * H, 8 * W) and r[4].shape == (8 * H, 8 * W)
        assert sorted(s['id'] for s in segs) == list(range(1, len(segs) + 1))
        assert set(np.unique(pan)) <= set([0] + [s['id'] for s in segs])


def test_two_frames_through_recycled_addresses(dev):
    """frame 2's feature maps come back from the caching allocator at frame 1's addresses with
    ``_version`` 0 again.  The prepared-feature cache must not mistake them for frame 1's (it compares object identity
    through weak references), so each frame's result equals the result of decoding that frame alone."""
    B, H, W, seed = 1, 16, 24, 0
    head = build_roi_head(dev, seed)
    frames = [synth.synth_decoder_inputs(B, H, W, s) for s in (0, 1)]
    want = []
    for f in frames:                        # each frame alone, fresh module state
        solo = build_roi_head(dev, seed)
        inp = {k: v.to(dev) for k, v in f.items()}
        with torch.no_grad():
            want.append(solo.decode(inp['x_feats'], inp['proposal_feats'], inp['mask_preds'], inp['depth_feats'],
                                    inp['depth_proposal'])['scaled_mask_preds'].clone())
        del inp, solo
    got, addrs = [], []
    for f in frames:                        # the frame loop of a video / dataset run: inputs freed between frames
        x, d = f['x_feats'].to(dev), f['depth_feats'].to(dev)
        addrs.append((x.data_ptr(), d.data_ptr()))
        with torch.no_grad():
            out = head.decode(x, f['proposal_feats'].to(dev), f['mask_preds'].to(dev), d, f['depth_proposal'].to(dev))
        got.append(out['scaled_mask_preds'].clone())
        torch.cuda.synchronize()
        del x, d, out
    # (the allocator normally hands the same blocks back; the assertion below holds either way)
    for i in range(2):
        assert torch.equal(got[i], want[i]), 'frame %d decoded from stale features (addresses %r)' % (i, addrs)
    assert not torch.equal(got[0], got[1])


def test_unselected_seeds_flip_count_and_effect(dev):
    """the un-teacher-forced loop on seeds that were NOT chosen for their margin from 0.  A mask
    logit within rounding of 0 may binarise differently than in the fp32 oracle; here the flips are COUNTED per stage and
    their effect on the final logits is bounded, instead of being avoided by seed selection."""
    from oracle import decoder_ref as ref
    from polyphonicformer_b200.decoder import DecoderEngine
    B, H, W = 1, 16, 24
    report = []
    for seed in range(4):
        sd = synth.synth_decoder_state(3, seed)
        stage_dicts = [{k[len('mask_head.%d.' % s):]: v for k, v in sd.items() if k.startswith('mask_head.%d.' % s)}
                       for s in range(3)]
        eng = DecoderEngine(stage_dicts, dev)
        inp = synth.synth_decoder_inputs(B, H, W, seed)
        with torch.no_grad():
            want = ref.decoder_forward(sd, inp['x_feats'], inp['proposal_feats'], inp['mask_preds'], inp['depth_feats'],
                                       inp['depth_proposal'], return_all_stages=True)
        feats = eng.prepare_feats(inp['x_feats'].to(dev), inp['depth_feats'].to(dev))
        mask = inp['mask_preds'].to(dev)
        obj = inp['proposal_feats'].reshape(B, -1, 256).to(dev)
        dep = inp['depth_proposal'].reshape(B, -1, 256).to(dev)
        nflip, nbits = [], mask.numel()
        for s in range(3):                  # feeding OUR outputs forward (no teacher forcing)
            cls, logits, obj, dep = eng.stage_forward(s, feats, mask, obj, dep, H, W)
            mask = logits[0]
            nflip.append(int(((mask.cpu() > 0) != (want['stages'][s]['mask_preds'] > 0)).sum()))
        l2, mx = rel_err(mask.cpu(), want['stages'][2]['mask_preds'])
        margin = min(float(st['mask_preds'].abs().min()) for st in want['stages'][:2])
        report.append((seed, nflip, margin, l2))
        # every flip of stages 0/1 moves one pooled row by one pixel's feature vector; with 384-pixel maps that is a
        # ~1e-2 perturbation of that row, so the final error may exceed the 1e-3 gate on such seeds -- but stays small
        assert all(n <= 1e-3 * nbits for n in nflip), report      # flips are rare (a few bits of 42 624) ...
        if nflip[0] == 0 and nflip[1] == 0:
            assert l2 < GATE, report                              # ... without one the loop meets the gate
        else:
            assert l2 < 0.2, report                               # ... with one the error is bounded, and reported
    print('seed, flipped mask bits per stage (of %d), min |logit| of stages 0/1, final rel err:' % nbits)
    for r in report:
        print('  ', r)
