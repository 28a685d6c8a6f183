"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol of include/pf_decoder.h, the
weight packing (feat_transform fold) is algebraically equal to the reference layers, and compute entry points fail
loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, rel_err
from oracle import decoder_ref as ref
from oracle import synth


@pytest.fixture(scope='module')
def lib():
    from polyphonicformer_b200 import build, _cabi
    build.build()
    return _cabi.load()


def test_library_exports_every_declared_symbol(lib):
    from polyphonicformer_b200 import _cabi
    inc = os.path.join(ROOT, 'include')
    hdr = ''.join(open(os.path.join(inc, f)).read() for f in sorted(os.listdir(inc)) if f.endswith('.h'))
    code = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'^(?:int|size_t|const char\*)\s+(pf_[a-z0-9_]+)\s*\(', code, flags=re.M))
    assert declared, 'no prototypes found'
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)
    assert lib.pf_version() == 100


def test_struct_layout_matches_header():
    from polyphonicformer_b200 import _cabi
    hdr = open(os.path.join(ROOT, 'include', 'pf_decoder.h')).read()
    body = hdr[hdr.index('typedef struct pf_branch_weights {'):hdr.index('} pf_branch_weights;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    ints = [n.strip() for decl in re.findall(r'^\s*int\s+([^;]+);', body, flags=re.M) for n in decl.split(',')]
    ptrs = [n.strip().lstrip('*') for decl in re.findall(r'const float\s*([^;]+);', body) for n in decl.split(',')]
    assert ints == _cabi.BranchWeights._ROWS + ['head_relu']
    assert ptrs == _cabi.BranchWeights._PTRS
    assert ctypes.sizeof(_cabi.BranchWeights) == 4 * 13 + 4 + 8 * len(ptrs)    # 13 ints, pad to 8, pointers
    assert ctypes.sizeof(_cabi.StageWeights) == 2 * ctypes.sizeof(_cabi.BranchWeights) + 16 + 16 + 8   # + vec_slices


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU behaviour')
def test_no_cpu_fallback(lib):
    from polyphonicformer_b200 import _cabi
    from polyphonicformer_b200.decoder import DecoderEngine
    buf = (ctypes.c_float * 64)()
    assert lib.pf_upsample2x(buf, buf, 1, 2, 2, None) == -4          # PF_ERR_ARCH, not a silent CPU result
    assert lib.pf_last_error_string()
    with pytest.raises(_cabi.PFError):
        DecoderEngine([], 'cpu')
    assert lib.pf_pool_splits(4, 2, 32768) == 18                       # host-only helper works without a device
    assert lib.pf_decoder_workspace_bytes(4, 111, 32768, 2048) > 0


def test_packed_weights_reproduce_reference_stage():
    """Replays the kernel launch plan of pf_kernel_update with torch ops on the PACKED (folded) weights and compares
    with the oracle stage -- proves the feat_transform fold and the layer wiring independent of any kernel."""
    from polyphonicformer_b200.decoder import PackedStage, gate_interleave_index
    import torch.nn.functional as F
    B, H, W, seed = 2, 12, 16, 3
    sd = synth.synth_decoder_state(1, seed)
    ssd = {k[len('mask_head.0.'):]: v for k, v in sd.items()}
    ps = PackedStage(ssd, 'cpu', synth.NUM_CLASSES, synth.FFN)
    inp = synth.synth_decoder_inputs(B, H, W, seed)
    N = inp['proposal_feats'].shape[1]
    x = [inp['x_feats'].reshape(B, 256, -1), inp['depth_feats'].reshape(B, 256, -1)]
    m = (inp['mask_preds'] > 0).float().reshape(B, N, -1)
    cnt = m.sum(-1)
    obj_in = inp['proposal_feats'].reshape(B, N, 256)
    dep_in = inp['depth_proposal'].reshape(B, N, 256)
    outs = {}

    def ln(t, p):
        return F.layer_norm(t, (256,), p[0], p[1], 1e-5)

    for bi in (0, 1):
        v = {k[1]: t for k, t in ps.views.items() if k[0] == bi}
        for name in PackedStage.MATS + ('ffn2_w',):   # the bf16 hi + lo planes reproduce the fp32 matrix to ~2^-17
            if name not in v:
                continue
            row = getattr(ps.struct.br[bi], name)
            stack = ps.packer.stack_ffn if name == 'ffn2_w' else ps.packer.stack256
            pad = (v[name].shape[0] + 127) // 128 * 128
            rec = stack[row:row + v[name].shape[0]].float() + stack[row + pad:row + pad + v[name].shape[0]].float()
            want_w = v[name][gate_interleave_index()] if name in PackedStage.PERMUTED else v[name]
            assert (rec - want_w).abs().max() <= 2.0 ** -16 * want_w.abs().max()
        pooled = torch.einsum('bnh,bch->bnc', m, x[bi])                       # raw features: no feat_transform
        params = pooled @ v['dyn_w'].t() + cnt[..., None] * v['dyn_cb'] + v['dyn_b']
        p_in, p_out = params[..., :256], ln(params[..., 256:], v['ln_norm_out'])
        inp_k = obj_in if bi == 0 else dep_in + obj_in
        iv = inp_k @ v['inp_w'].t() + v['inp_b']
        i_in, i_out = iv[..., :256], ln(iv[..., 256:], v['ln_input_norm_out'])
        gate = (i_in * p_in) @ v['gate_w'].t() + v['gate_b']
        ig = ln(gate[..., :256], v['ln_input_norm_in']).sigmoid()
        ug = ln(gate[..., 256:], v['ln_norm_in']).sigmoid()
        f = ug * p_out + ig * i_out
        o0 = torch.relu(ln(f @ v['fc_w'].t() + v['fc_b'], v['ln_fc_norm']))
        o1 = ln(ref.multihead_self_attention(
            {'attn.in_proj_weight': v['qkv_w'], 'attn.in_proj_bias': v['qkv_b'], 'attn.out_proj.weight': v['out_w'],
             'attn.out_proj.bias': v['out_b']}, o0.permute(1, 0, 2)).permute(1, 0, 2), v['ln_attn'])
        h = torch.relu(o1 @ v['ffn1_w'].t() + v['ffn1_b'])
        o2 = ln(o1 + h @ v['ffn2_w'].t() + v['ffn2_b'], v['ln_ffn'])
        head = o2 @ v['head_w'].t()
        if bi == 0:
            t_c = torch.relu(ln(head[..., :256], v['ln_head_a']))
            t_m = torch.relu(ln(head[..., 256:], v['ln_head_b']))
            outs['cls'] = (t_c @ v['cls_w'].t() + v['cls_b'])[..., :synth.NUM_CLASSES]
        else:
            t_m = ln(head, v['ln_head_a'])
        kern = t_m @ v['kern_w'].t() + v['kern_b']
        kbias = t_m @ v['kb_w'] + ps.kb_b[bi]
        outs[bi] = (o2, torch.einsum('bnc,bch->bnh', kern, x[bi]) + kbias[..., None])

    with torch.no_grad():
        cls, mask, obj, depth, dep = ref.kernel_update_head(ssd, inp['x_feats'], inp['proposal_feats'],
                                                            inp['mask_preds'], inp['depth_proposal'],
                                                            inp['depth_feats'])
    for got, want in ((outs['cls'], cls), (outs[0][1], mask.reshape(B, N, -1)), (outs[1][1], depth.reshape(B, N, -1)),
                      (outs[0][0], obj.reshape(B, N, 256)), (outs[1][0], dep.reshape(B, N, 256))):
        l2, mx = rel_err(got, want)
        assert l2 < 2e-5 and mx < 2e-5, (l2, mx)      # fp32 re-association of the fold only


def test_head_weights_struct_and_packing():
    """struct pf_head_weights layout and PackedKernelHead's row blocks (hi + lo planes reproduce the fp32 weights)."""
    from oracle import synth
    from polyphonicformer_b200 import _cabi
    from polyphonicformer_b200.kernel_head import PackedKernelHead, H_ROW_SEG, H_ROW_DEP
    hdr = open(os.path.join(ROOT, 'include', 'pf_decoder.h')).read()
    body = hdr[hdr.index('typedef struct pf_head_weights {'):hdr.index('} pf_head_weights;')]
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    names = re.findall(r'(\w+);', body)
    assert names == [f[0] for f in _cabi.HeadWeights._fields_]
    assert ctypes.sizeof(_cabi.HeadWeights) == 5 * 8 + 4 * 4
    sd = synth.synth_kernel_head_state(1)
    pk = PackedKernelHead(sd, 'cpu')
    conv = pk.conv_split.float().reshape(2, 3, 2, 128, 256)            # [half][map][plane][row][in]
    for m, name in enumerate(('loc', 'seg', 'depth')):
        w = sd[f'{name}_convs.0.conv.weight'].reshape(256, 256)
        rec = torch.cat([conv[0, m, 0] + conv[0, m, 1], conv[1, m, 0] + conv[1, m, 1]])
        assert (rec - w).abs().max() < 2e-5 * w.abs().max()
    hw = pk.head_w.float().sum(0)
    assert (hw[:100] - sd['init_kernels.weight'].reshape(100, 256)).abs().max() < 1e-5
    assert (hw[H_ROW_SEG:H_ROW_SEG + 19] - sd['conv_seg.weight'].reshape(19, 256)).abs().max() < 1e-5
    assert (hw[H_ROW_DEP] - sd['conv_direct_depth.weight'].reshape(256)).abs().max() < 1e-5
    assert float(hw[100:H_ROW_SEG].abs().max()) == 0 and float(hw[H_ROW_DEP + 1:].abs().max()) == 0
    assert torch.equal(pk.head_b[H_ROW_SEG:H_ROW_SEG + 19], sd['conv_seg.bias'])
    assert pk.stuff_kernels.shape == (11, 256)


def test_pinned_result_pool_lifetime(monkeypatch):
    """postprocess._PinnedPool: the buffer behind a result returns to the pool only when the LAST array derived from it is
    gone; a caller that keeps everything makes the pool stop lending (results then use ordinary pageable memory)."""
    import gc
    import numpy as np
    from polyphonicformer_b200 import postprocess as pp
    monkeypatch.setattr(torch.Tensor, 'pin_memory', lambda self: self)      # no CUDA here: plain memory stands in
    pool = pp._PinnedPool(max_outstanding=2)
    (a, ha), (b, hb) = pool.lend(3_000_000), pool.lend(100)
    assert pool.outstanding == 2 and pool.lend(5) is None
    arr = ha[:16].view(np.int32).reshape(2, 2)                              # what get_panoptic_batch hands out
    a[:16] = torch.arange(16, dtype=torch.uint8)
    assert arr[0, 0] == 0x03020100
    del a, ha
    gc.collect()
    assert pool.outstanding == 2                                            # `arr` still owns the buffer
    del arr
    gc.collect()
    assert pool.outstanding == 1 and len(pool.free[1 << 22]) == 1
    c = pool.lend(4_000_000)
    assert c is not None and pool.outstanding == 2 and not pool.free[1 << 22]


def test_track_and_fpn_struct_layouts_match_headers():
    """ctypes mirrors of the structs in include/pf_track.h and include/pf_fpn.h: same field names, order and sizes."""
    from polyphonicformer_b200 import _cabi

    def fields(header, name):
        hdr = open(os.path.join(ROOT, 'include', header)).read()
        body = hdr[hdr.index('typedef struct %s {' % name):hdr.index('} %s;' % name)]
        body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
        out = []
        for typ, decl in re.findall(r'^\s*(const uint16_t\*|const float\*|float|int)\s+([^;]+);', body, flags=re.M):
            out += [(n.strip().lstrip('*'), typ) for n in decl.split(',')]
        return out

    ctype = {'const uint16_t*': ctypes.c_void_p, 'const float*': ctypes.c_void_p, 'float': ctypes.c_float, 'int': ctypes.c_int}
    for header, name, cls in (('pf_track.h', 'pf_track_weights', _cabi.TrackWeights),
                              ('pf_track.h', 'pf_tracker_config', _cabi.TrackerConfig),
                              ('pf_fpn.h', 'pf_fpn_weights', _cabi.FpnWeights)):
        want = fields(header, name)
        assert want, name
        assert [(n, t) for n, t in cls._fields_] == [(n, ctype[t]) for n, t in want], name
    assert ctypes.sizeof(_cabi.TrackWeights) == 7 * 8 + 8 and ctypes.sizeof(_cabi.TrackerConfig) == 40
    assert ctypes.sizeof(_cabi.FpnWeights) == 3 * 8 + 8
