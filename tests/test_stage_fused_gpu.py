"""The fused cluster kernel of the small-N block (csrc/pf_stage.cu, one launch per stage) against the per-layer kernels
(csrc/pf_update.cu, 12 launches): same inputs, every output and every intermediate activation slot of the arena."""
import ctypes

import pytest
import torch

from conftest import rel_err
from oracle import synth

pytestmark = pytest.mark.gpu

SLOTS = ['POOLED', 'INP', 'GATEIN', 'MIX', 'OBJ0', 'ATT', 'OBJ1'] + ['HID%d' % i for i in range(8)] + ['OBJ2', 'HEAD0', 'HEAD1']


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    return torch.device('cuda:0')


def run_update(eng, stage, partial, cntp, S, obj, dep, B, N, fused, cls_sigmoid=0):
    from polyphonicformer_b200 import _cabi
    from polyphonicformer_b200.decoder import _ptr, _stream_ptr
    lib = _cabi.load()
    dev = obj.device
    old = lib.pf_set_fused_update(1 if fused else 0)
    try:
        nbytes = lib.pf_update_workspace_bytes(B, N, 2048)
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        out = dict(obj=torch.empty_like(obj), dep=torch.empty_like(dep),
                   cls=torch.empty((B, N, 19), dtype=torch.float32, device=dev),
                   kern=torch.empty((2 * B, N, 256), dtype=torch.float32, device=dev),
                   ksplit=torch.empty((2 * B, 2, N, 256), dtype=torch.bfloat16, device=dev),
                   kbias=torch.empty((2, B, N), dtype=torch.float32, device=dev))
        before = lib.pf_last_launch_count()
        _cabi.call('pf_kernel_update', ctypes.byref(eng.stages[stage].struct), _ptr(partial), _ptr(cntp), S, _ptr(obj),
                   _ptr(dep), _ptr(out['obj']), _ptr(out['dep']), _ptr(out['cls']), _ptr(out['kern']), _ptr(out['ksplit']),
                   _ptr(out['kbias']), _ptr(ws), nbytes, B, N, cls_sigmoid, _stream_ptr())
        torch.cuda.synchronize()
        out['launches'] = lib.pf_last_launch_count() - before
        arena = ws[:2 * B * len(SLOTS) * 2 * 128 * 256 * 2].view(torch.bfloat16)
        if fused:    # column-group-major blocks [32 groups][128 rows][8] per (unit, slot, plane) -> [128][256]
            arena = arena.view(2 * B, len(SLOTS), 2, 32, 128, 8).permute(0, 1, 2, 4, 3, 5).reshape(2 * B, len(SLOTS), 2, 128, 256)
        else:
            arena = arena.view(2 * B, len(SLOTS), 2, 128, 256)
        out['arena'] = (arena[:, :, 0].float() + arena[:, :, 1].float())[:, :, :N].clone()   # hi + lo, valid rows
        return out
    finally:
        lib.pf_set_fused_update(old)


def make_case(dev, B, H, W, N, seed):
    from polyphonicformer_b200 import _cabi
    from polyphonicformer_b200.decoder import DecoderEngine, _ptr, _stream_ptr
    lib = _cabi.load()
    sd = synth.synth_decoder_state(3, seed)
    stage_dicts = [{k[len('mask_head.%d.' % s):]: v for k, v in sd.items() if k.startswith('mask_head.%d.' % s)}
                   for s in range(3)]
    eng = DecoderEngine(stage_dicts, dev)
    inp = synth.synth_decoder_inputs(B, H, W, seed, n_kernels=N)
    feats = eng.prepare_feats(inp['x_feats'].to(dev), inp['depth_feats'].to(dev))
    HW, HWp = H * W, feats.shape[-1]
    S = lib.pf_pool_splits(B, 2, HW)
    bits = torch.empty((B, (HW + 31) // 32, 128), dtype=torch.int32, device=dev)
    partial = torch.empty((2 * B, S, N, 256), dtype=torch.float32, device=dev)
    cntp = torch.empty((2 * B, S, N), dtype=torch.float32, device=dev)
    mask = inp['mask_preds'].to(dev).contiguous()
    st = _stream_ptr()
    _cabi.call('pf_binarise', _ptr(mask), _ptr(bits), B, N, HW, st)
    _cabi.call('pf_mask_pool', _ptr(feats), _ptr(bits), _ptr(partial), _ptr(cntp), B, N, HW, HWp, 2, S, st)
    obj = inp['proposal_feats'].reshape(B, N, 256).to(dev).contiguous()
    dep = inp['depth_proposal'].reshape(B, N, 256).to(dev).contiguous()
    torch.cuda.synchronize()
    return eng, partial, cntp, S, obj, dep


@pytest.mark.parametrize('B,H,W,N', [(1, 16, 24, 111), (2, 16, 24, 111), (4, 32, 64, 111), (1, 10, 12, 5), (3, 8, 12, 128),
                                     (9, 8, 12, 111)])
def test_fused_stage_matches_per_layer_kernels(dev, B, H, W, N):
    eng, partial, cntp, S, obj, dep = make_case(dev, B, H, W, N, seed=3)
    for stage in (0, 2):
        want = run_update(eng, stage, partial, cntp, S, obj, dep, B, N, fused=False, cls_sigmoid=stage == 2)
        got = run_update(eng, stage, partial, cntp, S, obj, dep, B, N, fused=True, cls_sigmoid=stage == 2)
        assert want['launches'] == 12 and got['launches'] == 1
        report = []
        for i, name in enumerate(SLOTS):
            for br in range(2):
                a, b = got['arena'][br * B:(br + 1) * B, i], want['arena'][br * B:(br + 1) * B, i]
                if name == 'HEAD0' and br == 1:
                    continue                                  # the depth branch has no cls feature
                report.append((name, br) + rel_err(a.cpu(), b.cpu()))
        for k in ('obj', 'dep', 'cls', 'kern', 'kbias'):
            report.append((k, -1) + rel_err(got[k].cpu(), want[k].cpu()))
        report.append(('ksplit', -1) + rel_err(got['ksplit'].float().sum(1).cpu(), want['ksplit'].float().sum(1).cpu()))
        bad = [r for r in report if not (r[2] < 2e-5 and r[3] < 1e-4)]
        assert not bad, 'B=%d N=%d stage %d: first diverging tensors (name, branch, l2, max): %r\nall: %r' % (B, N, stage, bad[:4], report)


def test_fused_stage_in_place_and_deterministic(dev):
    """obj_out == obj_in (how pf_decoder_forward chains the stages) goes through the temporaries; reruns are bit-identical."""
    from polyphonicformer_b200 import _cabi
    from polyphonicformer_b200.decoder import _ptr, _stream_ptr
    lib = _cabi.load()
    B, N = 2, 111
    eng, partial, cntp, S, obj, dep = make_case(dev, B, 16, 24, N, seed=1)
    ref = run_update(eng, 1, partial, cntp, S, obj, dep, B, N, fused=True)
    again = run_update(eng, 1, partial, cntp, S, obj, dep, B, N, fused=True)
    for k in ('obj', 'dep', 'cls', 'kern', 'kbias', 'ksplit'):
        assert torch.equal(ref[k], again[k]), k
    o, d = obj.clone(), dep.clone()
    nbytes = lib.pf_update_workspace_bytes(B, N, 2048)
    ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    ks = torch.empty((2 * B, 2, N, 256), dtype=torch.bfloat16, device=dev)
    kb = torch.empty((2, B, N), dtype=torch.float32, device=dev)
    cls = torch.empty((B, N, 19), dtype=torch.float32, device=dev)
    old = lib.pf_set_fused_update(1)
    try:
        _cabi.call('pf_kernel_update', ctypes.byref(eng.stages[1].struct), _ptr(partial), _ptr(cntp), S, _ptr(o), _ptr(d), _ptr(o),
                   _ptr(d), _ptr(cls), None, _ptr(ks), _ptr(kb), _ptr(ws), nbytes, B, N, 0, _stream_ptr())
        torch.cuda.synchronize()
    finally:
        lib.pf_set_fused_update(old)
    assert torch.equal(o, ref['obj']), (o - ref['obj']).abs().max()
    assert torch.equal(d, ref['dep']), (d - ref['dep']).abs().max()
    assert torch.equal(ks, ref['ksplit'])
