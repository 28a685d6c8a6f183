"""GPU parity tests of the individual CUDA kernels, called through the C ABI (ctypes), against plain PyTorch
references computed on the same device.  Tolerances are written next to each comparison."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu

cabi = pytest.importorskip('polyphonicformer_b200._cabi')


@pytest.fixture(scope='module')
def dev():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    cabi.load()
    return torch.device('cuda:0')


def P(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def S():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def pack_bits_ref(mask_bool, N, HW):
    """[B,N,HW] bool -> [B,WORDS,128] int32 reference packing."""
    B = mask_bool.shape[0]
    words = (HW + 31) // 32
    m = torch.zeros((B, 128, words * 32), dtype=torch.int64, device=mask_bool.device)
    m[:, :N, :HW] = mask_bool.long()
    w = (m.view(B, 128, words, 32) << torch.arange(32, device=m.device)).sum(-1)   # [B,128,words]
    w = torch.where(w >= 2 ** 31, w - 2 ** 32, w)
    return w.permute(0, 2, 1).contiguous().to(torch.int32)


@pytest.mark.parametrize('B,HW', [(1, 120), (2, 384), (1, 2048)])
def test_cast_feats(dev, B, HW):
    torch.manual_seed(0)
    x = torch.randn(B, 256, HW, device=dev)
    d = torch.randn(B, 256, HW, device=dev)
    HWp = (HW + 7) // 8 * 8
    out = torch.full((2, B, 256, HWp), 7.0, dtype=torch.bfloat16, device=dev)
    cabi.call('pf_cast_feats', P(x), P(d), P(out), B, HW, HWp, S())
    torch.cuda.synchronize()
    assert torch.equal(out[0, :, :, :HW], x.to(torch.bfloat16))      # bit-exact: same RN rounding
    assert torch.equal(out[1, :, :, :HW], d.to(torch.bfloat16))
    assert (out[:, :, :, HW:] == 0).all()


@pytest.mark.parametrize('B,N,HW', [(1, 111, 120), (2, 111, 384), (1, 100, 2048), (1, 5, 33)])
def test_binarise(dev, B, N, HW):
    torch.manual_seed(1)
    logits = torch.randn(B, N, HW, device=dev)
    logits[0, 0, :5] = torch.tensor([0.0, -0.0, 1e-30, -1e-30, float('nan')], device=dev)
    words = (HW + 31) // 32
    bits = torch.full((B, words, 128), -1, dtype=torch.int32, device=dev)
    cabi.call('pf_binarise', P(logits), P(bits), B, N, HW, S())
    torch.cuda.synchronize()
    assert torch.equal(bits, pack_bits_ref(logits > 0, N, HW))       # bit-exact


@pytest.mark.parametrize('maps,H,W', [(3, 10, 12), (7, 16, 24), (2, 5, 7), (1, 1, 1), (4, 48, 156)])
def test_upsample2x(dev, maps, H, W):
    torch.manual_seed(2)
    x = torch.randn(1, maps, H, W, device=dev)
    out = torch.empty(1, maps, 2 * H, 2 * W, device=dev)
    cabi.call('pf_upsample2x', P(x), P(out), maps, H, W, S())
    torch.cuda.synchronize()
    ref = F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)
    assert torch.allclose(out, ref, rtol=0, atol=1e-6), (out - ref).abs().max()   # fp32 re-association only


@pytest.mark.parametrize('B,N,HW,nb', [(1, 111, 120, 2), (2, 111, 384, 2), (1, 100, 2048, 1), (1, 111, 7488, 2),
                                        (3, 111, 64, 2)])
def test_mask_pool(dev, B, N, HW, nb):
    torch.manual_seed(3)
    lib = cabi.load()
    HWp = (HW + 7) // 8 * 8
    feats = torch.randn(2, B, 256, HWp, device=dev).to(torch.bfloat16)
    mask = torch.rand(B, N, HW, device=dev) < 0.4
    bits = pack_bits_ref(mask, N, HW)
    Sp = lib.pf_pool_splits(B, nb, HW)
    assert Sp >= 1
    partial = torch.full((nb * B, Sp, N, 256), float('nan'), device=dev)
    cntp = torch.full((nb * B, Sp, N), float('nan'), device=dev)
    cabi.call('pf_mask_pool', P(feats), P(bits), P(partial), P(cntp), B, N, HW, HWp, nb, Sp, S())
    pooled = torch.empty(nb * B, N, 256, device=dev)
    count = torch.empty(B, N, device=dev)
    cabi.call('pf_pool_reduce', P(partial), P(cntp), P(pooled), P(count), B, N, nb, Sp, S())
    torch.cuda.synchronize()
    f = feats[:nb, :, :, :HW].double()
    ref = torch.einsum('bnh,gbch->gbnc', mask.double(), f).reshape(nb * B, N, 256)
    l2, mx = rel_err(pooled, ref)
    assert l2 < 2e-6 and mx < 2e-6, (l2, mx)     # products exact ({0,1} x bf16), fp32 accumulation order only
    assert torch.equal(count, mask.sum(-1).float())


@pytest.mark.parametrize('B,N,HW,units', [(1, 111, 120, 2), (2, 111, 384, 4), (2, 111, 384, 2), (1, 100, 2048, 2),
                                           (1, 111, 7488, 2)])
def test_mask_einsum(dev, B, N, HW, units):
    torch.manual_seed(4)
    HWp = (HW + 7) // 8 * 8
    feats = torch.randn(2, B, 256, HWp, device=dev).to(torch.bfloat16)
    kern = torch.randn(2, B, N, 256, device=dev) * 0.1
    kbias = torch.randn(2, B, N, device=dev)
    words = (HW + 31) // 32
    logits = torch.full((units, N, HW), float('nan'), device=dev)
    bits = torch.full((B, words, 128), -1, dtype=torch.int32, device=dev)
    ksplit = torch.full((2 * B, 2, N, 256), float('nan'), dtype=torch.bfloat16, device=dev)
    cabi.call('pf_split_kernels', P(kern), P(ksplit), 2 * B, N, S())
    torch.cuda.synchronize()
    hi = kern.reshape(2 * B, N, 256).to(torch.bfloat16)
    assert torch.equal(ksplit[:, 0], hi)
    assert torch.equal(ksplit[:, 1], (kern.reshape(2 * B, N, 256) - hi.float()).to(torch.bfloat16))
    cabi.call('pf_mask_einsum', P(feats), P(ksplit), P(kbias), P(logits), P(bits), B, N, HW, HWp, units, S())
    torch.cuda.synchronize()
    f = feats[:, :, :, :HW].double().reshape(2 * B, 256, HW)[:units]
    ref = torch.einsum('gnc,gch->gnh', kern.double().reshape(2 * B, N, 256)[:units], f) \
        + kbias.double().reshape(2 * B, N)[:units, :, None]
    l2, mx = rel_err(logits, ref)
    # hi/lo bf16 split of the kernel operand: ~2^-17 relative per product, fp32 accumulation
    assert l2 < 1e-5 and mx < 1e-5, (l2, mx)
    assert torch.equal(bits, pack_bits_ref(logits[:B] > 0, N, HW))   # bits are the sign of the emitted logits
    # bits-only mode must produce the same bits
    bits2 = torch.full_like(bits, -1)
    cabi.call('pf_mask_einsum', P(feats), P(ksplit), P(kbias), None, P(bits2), B, N, HW, HWp, B, S())
    torch.cuda.synchronize()
    assert torch.equal(bits2, bits)


def test_error_paths(dev):
    lib = cabi.load()
    x = torch.zeros(8, device=dev)
    assert lib.pf_upsample2x(None, P(x), 1, 1, 1, S()) == -1
    assert b'null' in lib.pf_last_error_string()
    with pytest.raises(cabi.PFError):
        cabi.call('pf_binarise', P(x), P(x), 1, 500, 8, S())        # N > 128
    assert lib.pf_mask_pool(P(x), P(x), P(x), P(x), 1, 111, 100, 100, 2, 1, S()) == -2   # pitch not multiple of 8
