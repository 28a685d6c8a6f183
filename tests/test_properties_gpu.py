"""Size-independent properties at BASELINE.json's FULL shapes (B=4, 1024x2048 frames -> 128x256 decoder map, N=111)
and the edge cases of the domain (empty / full masks, maximum and tiny N, ragged row groups), all through the C ABI.
The CPU oracle would need minutes at these sizes; the properties below have closed-form answers instead."""
import ctypes

import pytest
import torch

from conftest import rel_err
from oracle import decoder_ref as ref
from oracle import synth

pytestmark = pytest.mark.gpu

cabi = pytest.importorskip('polyphonicformer_b200._cabi')
B, H, W, N = 4, 128, 256, 111
HW = H * W


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


@pytest.fixture(scope='module')
def feats(dev):
    torch.manual_seed(7)
    return torch.randn(2, B, 256, HW, device=dev).to(torch.bfloat16)


def pool(feats, logits, n=N):
    lib = cabi.load()
    dev = feats.device
    words = (HW + 31) // 32
    bits = torch.empty((B, words, 128), dtype=torch.int32, device=dev)
    cabi.call('pf_binarise', P(logits), P(bits), B, n, HW, S())
    Sp = lib.pf_pool_splits(B, 2, HW)
    partial = torch.empty((2 * B, Sp, n, 256), device=dev)
    cntp = torch.empty((2 * B, Sp, n), device=dev)
    cabi.call('pf_mask_pool', P(feats), P(bits), P(partial), P(cntp), B, n, HW, HW, 2, Sp, S())
    pooled = torch.empty(2 * B, n, 256, device=dev)
    count = torch.empty(B, n, device=dev)
    cabi.call('pf_pool_reduce', P(partial), P(cntp), P(pooled), P(count), B, n, 2, Sp, S())
    torch.cuda.synchronize()
    return pooled.reshape(2, B, n, 256), count, bits


def test_pooling_full_size_closed_forms(dev, feats):
    """Row n < 37: empty mask -> exact zeros, count 0.  Row 37 <= n < 74: full mask -> the channel sums of the map.
    Rows >= 74: the left half of every image row -> the channel sums over that half; counts are exact integers."""
    logits = torch.empty(B, N, H, W, device=dev)
    logits[:, :37] = -1.0
    logits[:, 37:74] = 1.0
    logits[:, 74:] = -1.0
    logits[:, 74:, :, :W // 2] = 2.0
    pooled, count, _ = pool(feats, logits.reshape(B, N, HW))
    f = feats.double().reshape(2, B, 256, H, W)
    assert (pooled[:, :, :37] == 0).all() and (count[:, :37] == 0).all()
    full = f.sum((-1, -2))                                   # [2,B,256]
    l2, mx = rel_err(pooled[:, :, 37:74], full[:, :, None, :].expand(2, B, 37, 256))
    assert l2 < 2e-6 and mx < 2e-6, (l2, mx)
    half = f[..., :W // 2].sum((-1, -2))
    l2, mx = rel_err(pooled[:, :, 74:], half[:, :, None, :].expand(2, B, N - 74, 256))
    assert l2 < 2e-6 and mx < 2e-6, (l2, mx)
    assert (count[:, 37:74] == HW).all() and (count[:, 74:] == HW // 2).all()
    # all rows of a group see the same data: the split-K reduction must give bit-identical rows
    assert torch.equal(pooled[:, :, 37], pooled[:, :, 73]) and torch.equal(pooled[:, :, 74], pooled[:, :, N - 1])


def test_binarise_is_idempotent_on_its_own_output(dev):
    """bits(logits) == bits(2*bit - 1): thresholding the thresholded mask changes nothing (full size)."""
    torch.manual_seed(8)
    logits = torch.randn(B, N, HW, device=dev)
    logits[0, 0, :5] = torch.tensor([0.0, -0.0, 1e-38, -1e-38, 6e-8], device=dev)   # sigmoid(x) > 0.5 <=> x > 0
    words = HW // 32
    bits = torch.empty((B, words, 128), dtype=torch.int32, device=dev)
    cabi.call('pf_binarise', P(logits), P(bits), B, N, HW, S())
    again = (logits > 0).float() * 2 - 1
    bits2 = torch.empty_like(bits)
    cabi.call('pf_binarise', P(again), P(bits2), B, N, HW, S())
    torch.cuda.synchronize()
    assert torch.equal(bits, bits2)
    assert (bits[:, :, N:] == 0).all()                       # padding rows 111..127 stay clear
    first = bits[0, 0, 0].item() & 0x1F
    assert first == 0b10100                                  # 0, -0 -> 0; 1e-38 -> 1; -1e-38 -> 0; 6e-8 -> 1


def einsum(feats, kern, kbias, units=2 * B, want_bits=False):
    dev = feats.device
    ksplit = torch.empty((2 * B, 2, N, 256), dtype=torch.bfloat16, device=dev)
    cabi.call('pf_split_kernels', P(kern), P(ksplit), 2 * B, N, S())
    logits = torch.empty((units, N, HW), device=dev)
    bits = torch.empty((B, HW // 32, 128), dtype=torch.int32, device=dev) if want_bits else None
    cabi.call('pf_mask_einsum', P(feats), P(ksplit), P(kbias), P(logits), P(bits), B, N, HW, HW, units, S())
    torch.cuda.synchronize()
    return logits, bits


def test_einsum_full_size_selection_linearity_and_bits(dev, feats):
    """(i) a one-hot kernel row selects a feature channel EXACTLY (bf16 value, fp32 logit); (ii) the map is linear in
    the kernels: e(k1 + k2) = e(k1) + e(k2) up to fp32 rounding; (iii) the emitted bits are the sign of the logits."""
    torch.manual_seed(9)
    kern = torch.zeros(2, B, N, 256, device=dev)
    chan = torch.randint(0, 256, (2, B, N), device=dev)
    kern.scatter_(-1, chan[..., None], 1.0)
    zero_bias = torch.zeros(2, B, N, device=dev)
    logits, _ = einsum(feats, kern, zero_bias)
    want = torch.gather(feats.float(), 2, chan[..., None].expand(2, B, N, HW))
    assert torch.equal(logits.reshape(2, B, N, HW), want)
    k1 = torch.randn(2, B, N, 256, device=dev) * 0.1
    k2 = torch.randn(2, B, N, 256, device=dev) * 0.1
    b1, b2 = torch.randn(2, B, N, device=dev), torch.randn(2, B, N, device=dev)
    e1, _ = einsum(feats, k1, b1)
    e2, _ = einsum(feats, k2, b2)
    e12, bits = einsum(feats, k1 + k2, b1 + b2, want_bits=True)
    l2, mx = rel_err(e12, e1.double() + e2.double())
    assert l2 < 2e-5 and mx < 2e-5, (l2, mx)
    words = (e12[:B].reshape(B, N, HW // 32, 32) > 0).long()
    packed = (words << torch.arange(32, device=dev)).sum(-1)
    packed = torch.where(packed >= 2 ** 31, packed - 2 ** 32, packed).permute(0, 2, 1).to(torch.int32)
    assert torch.equal(bits[:, :, :N], packed)


def test_upsample_full_size_constant_and_ramp(dev):
    """Bilinear x2 (align_corners=False) keeps constants exactly and maps a horizontal ramp to the ramp sampled at
    (x' + 0.5)/2 - 0.5 in the interior; 2*B*N maps of 128x256 as in the last stage."""
    maps = 2 * B * N
    x = torch.empty(maps, H, W, device=dev)
    x[:] = torch.arange(maps, device=dev, dtype=torch.float32)[:, None, None]
    out = torch.empty(maps, 2 * H, 2 * W, device=dev)
    cabi.call('pf_upsample2x', P(x), P(out), maps, H, W, S())
    torch.cuda.synchronize()
    assert torch.equal(out, x[:, :1, :1].expand(maps, 2 * H, 2 * W))
    ramp = torch.arange(W, device=dev, dtype=torch.float32).expand(4, H, W).contiguous()
    out = torch.empty(4, 2 * H, 2 * W, device=dev)
    cabi.call('pf_upsample2x', P(ramp), P(out), 4, H, W, S())
    torch.cuda.synchronize()
    xs = ((torch.arange(2 * W, device=dev, dtype=torch.float32) + 0.5) / 2 - 0.5).clamp(0, W - 1)
    assert torch.allclose(out, xs.expand(4, 2 * H, 2 * W), rtol=0, atol=1e-5)


def make_engine(seed, dev):
    from polyphonicformer_b200.decoder import DecoderEngine
    sd = synth.synth_decoder_state(3, seed)
    stage_dicts = [{k[len('mask_head.%d.' % s):]: v for k, v in sd.items() if k.startswith('mask_head.%d.' % s)}
                   for s in range(3)]
    return DecoderEngine(stage_dicts, dev), sd


@pytest.mark.parametrize('n_kernels', [1, 5, 128])
def test_stage_with_extreme_kernel_counts_and_empty_masks(dev, n_kernels):
    """N = 1, 5 and the maximum 128 kernels; some masks entirely empty (count 0 -> pooled 0, only biases flow) and
    some entirely full, against the CPU oracle on the same inputs."""
    Bs, Hs, Ws, seed = 2, 12, 20, 4
    eng, sd = make_engine(seed, dev)
    inp = synth.synth_decoder_inputs(Bs, Hs, Ws, seed, n_kernels=n_kernels)
    inp['mask_preds'][:, 0] = -3.0                                  # empty mask
    if n_kernels > 2:
        inp['mask_preds'][:, 1] = 3.0                               # full mask
        inp['mask_preds'][0, 2] = -3.0
    ssd = {k[len('mask_head.0.'):]: v for k, v in sd.items() if k.startswith('mask_head.0.')}
    with torch.no_grad():
        cls, mask, obj, depth, dep = ref.kernel_update_head(ssd, inp['x_feats'], inp['proposal_feats'], inp['mask_preds'],
                                                            inp['depth_proposal'], inp['depth_feats'])
    feats = eng.prepare_feats(inp['x_feats'].to(dev), inp['depth_feats'].to(dev))
    g_cls, g_logits, g_obj, g_dep = eng.stage_forward(
        0, feats, inp['mask_preds'].to(dev), inp['proposal_feats'].reshape(Bs, n_kernels, 256).to(dev),
        inp['depth_proposal'].reshape(Bs, n_kernels, 256).to(dev), Hs, Ws)
    torch.cuda.synchronize()
    for got, want in ((g_cls, cls), (g_logits[0], mask), (g_logits[1], depth), (g_obj, obj.reshape(Bs, n_kernels, 256)),
                      (g_dep, dep.reshape(Bs, n_kernels, 256))):
        l2, mx = rel_err(got.cpu(), want)
        assert l2 < 5e-5 and mx < 5e-5, (n_kernels, l2, mx)


@pytest.mark.parametrize('rows', [1, 128, 300])
def test_kernel_updator_row_groups(dev, rows):
    """KernelUpdator over 1 row, exactly one 128-row group, and 2 full groups + a ragged tail of 44."""
    from polyphonicformer_b200.decoder import PackedUpdator, run_kernel_updator
    sd = {k[len('mask_head.0.kernel_update_conv.'):]: v for k, v in synth.synth_decoder_state(1, 0).items()
          if k.startswith('mask_head.0.kernel_update_conv.')}
    g = torch.Generator().manual_seed(rows)
    update = torch.randn(rows, 256, generator=g) * 20.0
    inputf = torch.randn(rows, 256, generator=g)
    with torch.no_grad():
        want = ref.kernel_updator(sd, update, inputf.reshape(rows, 1, 256)).reshape(rows, 256)
    out = run_kernel_updator(PackedUpdator(sd, dev), update.to(dev), inputf.to(dev))
    torch.cuda.synchronize()
    l2, mx = rel_err(out.cpu(), want)
    assert l2 < 5e-5 and mx < 5e-5, (rows, l2, mx)


def test_init_proposals_matches_kernel_head_tail(dev):
    """KernelHead._decode_init_proposals' tail (kernel_head.py:313-336) restated with plain PyTorch ops vs
    DecoderEngine.init_proposals (binarise + one-branch pooling + pf_init_proposals)."""
    Bs, Hs, Ws, P, T, S_ = 2, 24, 40, synth.N_PROPOSALS, synth.NUM_THING, synth.NUM_STUFF
    eng, _ = make_engine(0, dev)
    g = torch.Generator().manual_seed(11)
    x = synth.bf16_round(torch.relu(torch.randn(Bs, 256, Hs, Ws, generator=g)))
    d = synth.bf16_round(torch.relu(torch.randn(Bs, 256, Hs, Ws, generator=g)))
    mask = torch.randn(Bs, P, Hs, Ws, generator=g) - 0.3
    mask[0, 3] = -2.0                                                     # an empty mask: obj_feats = 0
    seg = torch.randn(Bs, T + S_, Hs, Ws, generator=g)
    w_init = torch.randn(P, 256, 1, 1, generator=g)
    w_seg = torch.randn(T + S_, 256, 1, 1, generator=g) * 0.01
    w_dd = torch.randn(1, 256, 1, 1, generator=g) * 0.01
    # kernel_head.py:313-336 (use_binary=True, proposal_feats_with_obj=True, cat_stuff_mask=True, eval)
    obj = torch.einsum('bnhw,bchw->bnc', (mask.sigmoid() > 0.5).float(), x)
    want_prop = torch.cat([w_init[None].expand(Bs, -1, -1, -1, -1) + obj.view(Bs, P, 256, 1, 1),
                           w_seg[T:][None].expand(Bs, -1, -1, -1, -1)], dim=1)
    want_mask = torch.cat([mask, seg[:, T:]], dim=1)
    feats = eng.prepare_feats(x.to(dev), d.to(dev))
    prop, masks, dprop = eng.init_proposals(feats, mask.to(dev), w_init.to(dev), seg.to(dev), w_seg.to(dev), w_dd.to(dev), T)
    torch.cuda.synchronize()
    assert prop.shape == (Bs, P + S_, 256, 1, 1) and dprop.shape == (Bs, P + S_, 256, 1, 1)
    l2, mx = rel_err(prop.cpu(), want_prop)
    assert l2 < 2e-6 and mx < 2e-6, (l2, mx)
    assert torch.equal(prop[:, P:].cpu(), want_prop[:, P:]) and torch.equal(masks.cpu(), want_mask)
    assert torch.equal(prop[0, 3].cpu().flatten(), w_init[3].flatten())   # empty mask -> the bare init kernel
    assert torch.equal(dprop[1, 7].cpu(), w_dd[0])


def test_full_size_decode_is_bit_reproducible_under_graph_replay(dev):
    """B=4 at 128x256: 25 replays of the captured step (every kernel launched with programmatic dependent launch,
    feature tiles prefetched ahead of the grid dependency) give bit-identical outputs -- a race between overlapping
    kernels of the chain would show up here."""
    eng, _ = make_engine(0, dev)
    g = torch.Generator().manual_seed(5)
    x = torch.relu(torch.randn(B, 256, H, W, generator=g)).to(torch.bfloat16).to(dev)
    d = torch.relu(torch.randn(B, 256, H, W, generator=g)).to(torch.bfloat16).to(dev)
    mask = (torch.randn(B, N, H, W, generator=g) - 0.3).to(dev)
    prop = (torch.randn(B, N, 256, generator=g) * 0.5).to(dev)
    dprop = (torch.randn(B, N, 256, generator=g) * 0.1).to(dev)
    feats = eng.prepare_feats(x, d)
    buf = eng.alloc_decode_buffers(B, N, H, W)

    def step():
        buf['obj'].copy_(prop), buf['dep'].copy_(dprop)
        eng.decode_inplace(feats, mask, buf, H, W)

    graph = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream(dev)
    cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cap):
        step()
        cap.synchronize()
        with torch.cuda.graph(graph, stream=cap):
            step()
    torch.cuda.current_stream().wait_stream(cap)
    graph.replay()
    torch.cuda.synchronize()
    ref_out = {k: buf[k].clone() for k in ('scaled', 'logits', 'cls', 'obj', 'dep')}
    assert torch.isfinite(ref_out['scaled']).all()
    for _ in range(25):
        graph.replay()
    torch.cuda.synchronize()
    for k, v in ref_out.items():
        assert torch.equal(buf[k], v), k
    step()                                                   # eager launches agree with the replayed graph
    torch.cuda.synchronize()
    for k, v in ref_out.items():
        assert torch.equal(buf[k], v), k
