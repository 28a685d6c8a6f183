"""Pins oracle/decoder_ref.py (the CPU restatement) against outputs of the
reference's own code (tests/golden/*.npz, made by oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import decoder_ref as ref
from oracle import synth
from conftest import rel_err, GOLDEN

CASES = ['decoder_b2_h16_w24_s6', 'decoder_b1_h10_w12_s1']
TOL = 2e-5   # fp32 re-association between the reference's op order and the restatement


@pytest.mark.parametrize('name', CASES)
def test_decoder_restatement_matches_reference(name):
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    B, H, W, seed = int(g['B']), int(g['H']), int(g['W']), int(g['seed'])
    sd = synth.synth_decoder_state(3, seed)
    inp = synth.synth_decoder_inputs(B, H, W, seed)
    with torch.no_grad():
        out = ref.decoder_forward(sd, inp['x_feats'], inp['proposal_feats'], inp['mask_preds'],
                                  inp['depth_feats'], inp['depth_proposal'], return_all_stages=True)
    for s, st in enumerate(out['stages']):
        for k, v in st.items():
            l2, mx = rel_err(v, g[f's{s}.{k}'])
            assert l2 < TOL and mx < TOL, (name, s, k, l2, mx)
        flips = (st['mask_preds'] > 0) != (torch.from_numpy(g[f's{s}.mask_preds']) > 0)
        assert flips.float().mean().item() < 1e-4
    for k, gk in (('scaled_mask_preds', 'scaled_mask_preds'), ('scaled_depth_preds', 'scaled_depth_preds'),
                  ('cls_score', 'cls_score_sigmoid')):
        l2, mx = rel_err(out[k], g[gk])
        assert l2 < TOL and mx < TOL, (name, k, l2, mx)


def test_updator_restatement_matches_reference():
    g = np.load(os.path.join(GOLDEN, 'updator_r37_s0.npz'))
    sd = {k[len('mask_head.0.kernel_update_conv.'):]: v
          for k, v in synth.synth_decoder_state(1, 0).items()
          if k.startswith('mask_head.0.kernel_update_conv.')}
    with torch.no_grad():
        y = ref.kernel_updator(sd, torch.from_numpy(g['update_feature']), torch.from_numpy(g['input_feature']))
    l2, mx = rel_err(y, g['out'])
    assert l2 < TOL and mx < TOL, (l2, mx)


def test_state_table_matches_reference_layout():
    """4.02 M parameters per stage."""
    n = sum(int(np.prod(s)) for s in synth.stage_state_shapes().values())
    assert abs(n / 1e6 - 4.02) < 0.01, n


PANOPTIC_CASES = ['panoptic_h32_w64_s0', 'panoptic_h24_w40_crop_s1']


def panoptic_args(g):
    """(roi_head stand-in, last_head stand-in, test_cfg, img_meta, inputs) of a panoptic golden case."""
    import json
    from types import SimpleNamespace
    from polyphonicformer_b200.registry import to_config
    h, w, seed = int(g['h']), int(g['w']), int(g['seed'])
    H0, W0 = [int(v) for v in g['img_hw']]
    cfg = to_config(json.load(open(os.path.join(GOLDEN, 'roi_head_cfg.json')))['test_cfg'])
    roi = SimpleNamespace(num_proposals=synth.N_PROPOSALS, num_thing_classes=synth.NUM_THING, merge_joint=True)
    last = SimpleNamespace(depth_act_mode='sigmoid', num_classes=synth.NUM_CLASSES)   # the shipped config (roi_head_cfg.json)
    meta = dict(img_shape=(H0, W0, 3), ori_shape=(H0, W0, 3), pad_shape=(4 * h, 4 * w, 3), scale_factor=1.0, flip=False,
                batch_input_shape=(4 * h, 4 * w))
    return roi, last, cfg, meta, synth.synth_panoptic_inputs(h, w, seed)


def segments_as_array(info):
    return np.array([[s['id'], int(s['isthing']), s['category_id'], s.get('instance_id', -1), s.get('area', -1)]
                     for s in info], dtype=np.int64).reshape(-1, 5)


@pytest.mark.parametrize('name', PANOPTIC_CASES)
def test_panoptic_restatement_matches_reference(name):
    """oracle/panoptic_ref.py == the reference's get_panoptic (kernel_update.py:421-535) on the same CPU ops: exact."""
    from oracle import panoptic_ref
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    roi, last, cfg, meta, inp = panoptic_args(g)
    with torch.no_grad():
        _, _, (pan, info), dbasic, dfinal = panoptic_ref.get_panoptic(
            roi, last, inp['cls_scores'], inp['mask_preds'], cfg, meta, inp['depth_preds'], inp['depth_init'])
    assert np.array_equal(pan, g['panoptic'])
    assert np.array_equal(segments_as_array(info), g['seg'])
    assert np.array_equal(dbasic, g['depth_basic']) and np.array_equal(dfinal, g['depth_final'])


@pytest.mark.parametrize('name', ['kernel_head_b2_h16_w24', 'kernel_head_b1_h10_w13'])
def test_kernel_head_restatement_matches_reference(name):
    """oracle/kernel_head_ref.py against the real KernelHead._decode_init_proposals (kernel_head.py:240-347)."""
    from oracle import kernel_head_ref
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    B, H, W, seed = int(g['B']), int(g['H']), int(g['W']), int(g['seed'])
    with torch.no_grad():
        out = kernel_head_ref.decode_init_proposals(synth.synth_kernel_head_state(seed), synth.synth_fpn_maps(B, H, W, seed))
    for k in ('proposal_feats', 'x_feats', 'mask_preds', 'seg_preds', 'depth_feats', 'depth_proposal', 'depth_pred'):
        l2, mx = rel_err(out[k], g[k])
        assert l2 < 1e-6 and mx < 1e-6, (name, k, l2, mx)
    assert float(g['margin']) > 5e-5


@pytest.mark.parametrize('name', ['fpn_pred_b2_h16_w24_s0', 'fpn_pred_b1_h10_w13_s1'])
def test_fpn_pred_restatement_matches_reference(name):
    """oracle/kernel_head_ref.fpn_pred against the real SemanticFPNWrapper conv_pred / aux_convs (semantic_fpn.py:221-229)."""
    from oracle import kernel_head_ref
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    B, H, W, seed = int(g['B']), int(g['H']), int(g['W']), int(g['seed'])
    with torch.no_grad():
        out = torch.stack(kernel_head_ref.fpn_pred(synth.synth_fpn_pred_state(seed), synth.synth_fused_map(B, H, W, seed)))
    l2, mx = rel_err(out, g['maps'])
    assert l2 < 1e-6 and mx < 1e-6, (name, l2, mx)


def test_semantic_fpn_restatement_matches_reference():
    """oracle/semantic_fpn_ref.py (the oracle of the next row) against the real
    SemanticFPNWrapper.forward (semantic_fpn.py:198-235): sine positional encoding, the 3x3 conv + GN + ReLU pyramid
    with its bilinear x2 steps, sum fusion, conv_pred + aux_convs."""
    from oracle import semantic_fpn_ref
    name = 'semantic_fpn_b1_h16_w24_s0'
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    B, H, W, seed = int(g['B']), int(g['H']), int(g['W']), int(g['seed'])
    with torch.no_grad():
        _, maps = semantic_fpn_ref.semantic_fpn_forward(synth.synth_semantic_fpn_state(seed),
                                                        synth.synth_fpn_inputs(B, H, W, seed))
    l2, mx = rel_err(torch.stack(maps), g['maps'])
    assert l2 < 1e-6 and mx < 1e-6, (l2, mx)


def test_tracking_restatement_matches_reference():
    """oracle/tracking_ref.py (the oracle of a row that is not built yet) against the reference's own
    SingleRoIExtractor + RoIAlign, QuasiDenseMaskEmbedHeadGTMask, mask -> box helpers and QuasiDenseEmbedTracker on a
    synthetic 4-frame clip (polyphonic_former_video.py:364-419): RoIs, embeddings, surviving boxes and track ids."""
    from oracle import tracking_ref
    g = np.load(os.path.join(GOLDEN, 'tracking_clip_s0.npz'))
    sd = synth.synth_track_head_state(int(g['seed']))
    tracker = tracking_ref.QuasiDenseTracker()
    seen = set()
    with torch.no_grad():
        for t, fr in enumerate(synth.synth_clip(seed=int(g['seed']))):
            fm = fr['masks'].float()
            boxes = torch.stack([tracking_ref.roi_box_of_mask(m) for m in fm])
            rois = torch.cat([boxes.new_zeros((len(boxes), 1)), boxes], 1).clamp(min=0)
            assert np.allclose(rois.numpy(), g[f'f{t}.rois'], rtol=0, atol=1e-5)
            emb = tracking_ref.track_forward(sd, fr['feats'], fm)
            l2, mx = rel_err(emb, g[f'f{t}.embeds'])
            assert l2 < 1e-6 and mx < 1e-6, (t, l2, mx)
            ids, kept = tracking_ref.track_frame(sd, tracker, fr['feats'], fr['masks'], fr['labels'], fr['scores'], t + 1)
            assert ids.tolist() == g[f'f{t}.ids'].tolist(), (t, ids.tolist(), g[f'f{t}.ids'].tolist())
            assert np.allclose(kept.numpy(), g[f'f{t}.boxes'], rtol=0, atol=1e-5)
            seen.update(ids.tolist())
    assert 0 in seen and len(seen) >= 6     # backdrops, persistent tracks and new tracks all occur in the clip
