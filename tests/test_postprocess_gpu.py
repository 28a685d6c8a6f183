"""GPU parity of pf_panoptic (through the host wrapper's C-ABI call) against the REAL reference's get_panoptic
(tests/golden/panoptic_*.npz) and, at the full 1024x2048 size, against the PyTorch restatement run on the device.

The panoptic map is an argmax over 111 products score * bilinear(sigmoid(logit)): where two products agree to within
fp32 rounding, any two implementations (the reference on CPU vs the reference on CUDA, for that matter) may pick a
different winner.  So integer parity is asserted as: identical segment list, areas within 0.2 %, and at most 0.1 % of
the pixels different -- and every differing pixel must be such a near-tie."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import panoptic_ref, synth
from test_oracle_golden import PANOPTIC_CASES, panoptic_args, segments_as_array

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    return torch.device('cuda:0')


def run(dev, roi, last, cfg, meta, inp):
    from polyphonicformer_b200 import postprocess
    d = {k: v.to(dev) for k, v in inp.items()}
    out = postprocess.get_panoptic(roi, last, d['cls_scores'], d['mask_preds'], cfg, meta, d['depth_preds'], d['depth_init'])
    torch.cuda.synchronize()
    return out


def near_tie_fraction(inp, meta, pan_a, pan_b, cfg, roi):
    """Of the pixels where the two maps differ: the fraction whose top-2 products differ by < 1e-5 relative."""
    diff = pan_a != pan_b
    if not diff.any():
        return 1.0
    P, T = roi.num_proposals, roi.num_thing_classes
    cls = inp['cls_scores']
    ts, idx = cls[:P, :T].flatten().topk(cfg.max_per_img)
    ss, sidx = cls[P:, T:].diag().sort(descending=True)
    masks = torch.cat([inp['mask_preds'][:P][idx // T], inp['mask_preds'][P:][sidx]])
    up = panoptic_ref.rescale_masks(masks, meta)
    prob = torch.cat([ts, ss]).view(-1, 1, 1) * up
    top2 = prob.topk(2, dim=0).values
    gap = ((top2[0] - top2[1]) / top2[0].clamp_min(1e-30))[torch.from_numpy(diff)]
    return (gap < 1e-5).float().mean().item()


@pytest.mark.parametrize('name', PANOPTIC_CASES)
def test_panoptic_matches_reference_golden(dev, name):
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    roi, last, cfg, meta, inp = panoptic_args(g)
    _, _, (pan, info), dbasic, dfinal = run(dev, roi, last, cfg, meta, inp)
    assert pan.dtype == np.int32 and pan.shape == g['panoptic'].shape
    got, want = segments_as_array(info), g['seg']
    assert np.array_equal(got[:, :4], want[:, :4])                       # ids, isthing, category, instance id
    stuff = want[:, 4] >= 0
    assert np.all(np.abs(got[stuff, 4] - want[stuff, 4]) <= np.maximum(2, 0.002 * want[stuff, 4]))
    scores = np.array([s.get('score', -1.0) for s in info])
    assert np.allclose(scores, g['seg_score'], rtol=0, atol=1e-7)
    ndiff, npix = int((pan != g['panoptic']).sum()), pan.size
    ties = near_tie_fraction(inp, meta, pan, g['panoptic'], cfg, roi)
    # first-max tie-breaking (lowest entry index wins an exact tie, as torch.argmax does) is what pp_argmax_kernel
    # implements; what remains are products that differ in the last bits between ATen-CPU and device arithmetic
    print('%s: %d of %d panoptic pixels differ from the reference (%.4f %%), %.0f %% of them near-ties (top-2 within 1e-5)'
          % (name, ndiff, npix, 100.0 * ndiff / npix, 100 * ties))
    assert ndiff <= 1e-3 * npix, (ndiff, npix)
    assert ties == 1.0, (ndiff, ties)
    same = pan == g['panoptic']
    assert np.allclose(dbasic, g['depth_basic'], rtol=1e-5, atol=1e-6)
    assert np.allclose(dfinal[same], g['depth_final'][same], rtol=1e-5, atol=1e-6)


def test_panoptic_full_size_against_restatement_on_device(dev):
    """1024x2048 frame (scaled predictions 256x512, N = 111): pf_panoptic vs oracle/panoptic_ref.py run with CUDA
    PyTorch ops on the same inputs."""
    from types import SimpleNamespace
    h, w = 256, 512
    g = dict(h=h, w=w, seed=3, img_hw=np.array([1024, 2048]))
    roi, last, cfg, meta, inp = panoptic_args(g)
    _, _, (pan, info), dbasic, dfinal = run(dev, roi, last, cfg, meta, inp)
    d = {k: v.to(dev) for k, v in inp.items()}
    with torch.no_grad():
        _, _, (pan_r, info_r), dbasic_r, dfinal_r = panoptic_ref.get_panoptic(
            roi, last, d['cls_scores'], d['mask_preds'], cfg, meta, d['depth_preds'], d['depth_init'])
    assert len(info) == len(info_r) >= 5
    got, want = segments_as_array(info), segments_as_array(info_r)
    assert np.array_equal(got[:, :4], want[:, :4])
    print('full size: %d of %d panoptic pixels differ from the restatement on the device' % (int((pan != pan_r).sum()), pan.size))
    assert (pan != pan_r).mean() <= 1e-4
    same = pan == pan_r
    assert np.allclose(dbasic, dbasic_r, rtol=1e-5, atol=1e-6)
    assert np.allclose(dfinal[same], dfinal_r[same], rtol=1e-5, atol=1e-6)
    assert set(np.unique(pan)) == set(range(len(info) + 1))


def test_panoptic_unsupported_geometry_raises(dev):
    g = np.load(os.path.join(GOLDEN, PANOPTIC_CASES[0] + '.npz'))
    roi, last, cfg, meta, inp = panoptic_args(g)
    meta = dict(meta, ori_shape=(200, 400, 3))                            # a real rescale: not covered, must not fall back
    with pytest.raises(NotImplementedError):
        run(dev, roi, last, cfg, meta, inp)


def test_panoptic_pipeline_matches_module_simple_test(dev):
    """PanopticPipeline (host buffers, overlapped submissions) == decode + per-frame get_panoptic done step by step."""
    from types import SimpleNamespace
    import json
    from polyphonicformer_b200 import postprocess
    from polyphonicformer_b200.decoder import DecoderEngine, PanopticPipeline
    from polyphonicformer_b200.registry import to_config
    B, H, W = 2, 16, 24
    sd = synth.synth_decoder_state(3, 0)
    stage_dicts = [{k[len('mask_head.%d.' % s):]: v for k, v in sd.items() if k.startswith('mask_head.%d.' % s)}
                   for s in range(3)]
    eng = DecoderEngine(stage_dicts, dev)
    cfg = to_config(json.load(open(os.path.join(GOLDEN, 'roi_head_cfg.json')))['test_cfg'])
    roi = SimpleNamespace(num_proposals=synth.N_PROPOSALS, num_thing_classes=synth.NUM_THING, merge_joint=True)
    last = SimpleNamespace(depth_act_mode='sigmoid', num_classes=synth.NUM_CLASSES)
    meta = dict(img_shape=(8 * H, 8 * W, 3), ori_shape=(8 * H, 8 * W, 3), batch_input_shape=(8 * H, 8 * W))
    pipe = PanopticPipeline(eng, B, synth.N_KERNELS, H, W)
    ins, outs, want = [], [], []
    for seed in range(3):
        inp = synth.synth_decoder_inputs(B, H, W, seed)
        feats = eng.prepare_feats(inp['x_feats'].to(dev), inp['depth_feats'].to(dev))
        o = eng.decode(feats, inp['mask_preds'].to(dev), inp['proposal_feats'].to(dev), inp['depth_proposal'].to(dev), H, W)
        dinit = eng.upsample2x(inp['depth_pred'].to(dev))
        want.append([postprocess.get_panoptic(roi, last, o['cls_score'][b], o['scaled_mask_preds'][b], cfg, meta,
                                              o['scaled_depth_preds'][b], dinit[b]) for b in range(B)])
        ins.append(dict(x=inp['x_feats'].to(torch.bfloat16).pin_memory(), d=inp['depth_feats'].to(torch.bfloat16).pin_memory(),
                        mask=inp['mask_preds'].pin_memory(), prop=inp['proposal_feats'].reshape(B, -1, 256).pin_memory(),
                        dprop=inp['depth_proposal'].reshape(B, -1, 256).contiguous().pin_memory(),
                        depth_pred=inp['depth_pred'].pin_memory()))
        outs.append(dict(panoptic=torch.empty((B, 8 * H, 8 * W), dtype=torch.int32).pin_memory(),
                         depth_final=torch.empty((B, 8 * H, 8 * W)).pin_memory(),
                         depth_basic=torch.empty((B, 8 * H, 8 * W)).pin_memory(),
                         segments=torch.empty((B, 128, 24), dtype=torch.uint8).pin_memory(),
                         nseg=torch.empty(B, dtype=torch.int32).pin_memory()))
    for i in range(3):
        pipe.submit(ins[i], outs[i])
    pipe.drain()
    for i in range(3):
        for b in range(B):
            _, _, (pan, info), dbasic, dfinal = want[i][b]
            assert np.array_equal(outs[i]['panoptic'][b].numpy(), pan)
            assert np.array_equal(outs[i]['depth_final'][b].numpy(), dfinal)
            assert np.array_equal(outs[i]['depth_basic'][b].numpy(), dbasic)
            assert int(outs[i]['nseg'][b]) == len(info)


def test_panoptic_batch_from_stride8_maps_is_bit_identical(dev):
    """pf_panoptic_batch(in_stride2=1) on the decoder's own stride-8 maps == pf_upsample2x then pf_panoptic per frame
    (kernel_update.py:131-143, 302-307 -> :421-469), for a batch whose frames also carry different crops."""
    from polyphonicformer_b200 import postprocess
    from polyphonicformer_b200.decoder import DecoderEngine
    B, h2, w2 = 3, 24, 40                                      # stride-8 maps; scaled predictions 48 x 80
    g = dict(h=2 * h2, w=2 * w2, seed=5, img_hw=np.array([8 * h2, 8 * w2]))
    roi, last, cfg, meta, _ = panoptic_args(g)
    frames = [synth.synth_panoptic_inputs(h2, w2, seed) for seed in (5, 6, 7)]     # blobs / bands at stride 8
    cls = torch.stack([f['cls_scores'] for f in frames]).to(dev)
    mask = torch.stack([f['mask_preds'] for f in frames]).to(dev)
    depth = torch.stack([f['depth_preds'] for f in frames]).to(dev)
    dinit = torch.stack([f['depth_init'] for f in frames]).to(dev)                # [B,1,h2,w2]
    eng = SimpleNamespaceEngine(dev)
    mask_up, depth_up, dinit_up = (eng.upsample2x(t) for t in (mask, depth, dinit))
    metas = [dict(meta), dict(meta, img_shape=(8 * h2 - 5, 8 * w2 - 9, 3), ori_shape=(8 * h2 - 5, 8 * w2 - 9, 3)), dict(meta)]
    got = postprocess.get_panoptic_batch(roi, last, cls, mask, cfg, metas, depth, dinit, stride2_inputs=True)
    assert len(got) == B
    for b in range(B):
        want = postprocess.get_panoptic(roi, last, cls[b], mask_up[b], cfg, metas[b], depth_up[b], dinit_up[b, 0])
        assert np.array_equal(got[b][2][0], want[2][0])
        assert got[b][2][1] == want[2][1] and len(want[2][1]) >= 2
        assert np.array_equal(got[b][3], want[3]) and np.array_equal(got[b][4], want[4])
    # two frames of one shape in one call == the same frames one by one
    pair = postprocess.get_panoptic_batch(roi, last, cls[::2], mask[::2], cfg, [metas[0], metas[2]], depth[::2], dinit[::2],
                                          stride2_inputs=True)
    for j, b in enumerate((0, 2)):
        assert np.array_equal(pair[j][2][0], got[b][2][0]) and pair[j][2][1] == got[b][2][1]
        assert np.array_equal(pair[j][4], got[b][4])


class SimpleNamespaceEngine:
    """pf_upsample2x without a DecoderEngine (no weights needed)."""

    def __init__(self, dev):
        self.dev = dev

    def upsample2x(self, maps):
        import ctypes
        from polyphonicformer_b200 import _cabi
        maps = maps.float().contiguous()
        H, W = maps.shape[-2:]
        out = torch.empty(maps.shape[:-2] + (2 * H, 2 * W), dtype=torch.float32, device=maps.device)
        _cabi.call('pf_upsample2x', ctypes.c_void_p(maps.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                   maps.numel() // (H * W), H, W, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        return out
