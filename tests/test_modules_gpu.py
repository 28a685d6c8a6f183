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
