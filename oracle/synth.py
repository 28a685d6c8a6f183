"""TEST INFRASTRUCTURE ONLY -- deterministic synthetic weights and decoder inputs.

Both ``oracle/make_golden.py`` (which runs the real reference here) and the
tests (which run on a box where /root/reference does not exist) call these
functions, so the golden files only need to hold the reference's OUTPUTS: the
inputs and the ``state_dict`` are regenerated bit-identically from key names,
shapes and a seed with torch's CPU generator.

Weights are deliberately *not* the reference's ``init_weights()`` values: biases
and LayerNorm affine parameters are made non-trivial so that a kernel that drops
a bias or a gamma cannot pass parity by accident.
"""
import math
import zlib

import torch

NUM_THING, NUM_STUFF = 8, 11
NUM_CLASSES = NUM_THING + NUM_STUFF
C = 256
FFN = 2048
HEADS = 8
N_PROPOSALS = 100
N_KERNELS = N_PROPOSALS + NUM_STUFF   # 111 at inference (kernel_head.py:329-336)


def _gen(key, seed):
    g = torch.Generator(device='cpu')
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def synth_tensor(key, shape, seed=0):
    """Deterministic tensor for a state-dict key (xavier-uniform-like matrices,
    gamma ~ 1 +- 0.1, other vectors ~ 0.1 * N(0,1))."""
    g = _gen(key, seed)
    shape = tuple(shape)
    if len(shape) >= 2:
        fan_out, fan_in = shape[0], int(math.prod(shape[1:]))
        a = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(shape, generator=g) * 2 - 1) * a
    if key.endswith('weight'):          # 1-D weight == LayerNorm / GroupNorm gamma
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    return 0.1 * torch.randn(shape, generator=g)


def stage_state_shapes():
    """name -> shape for one KernelUpdateHead built from
    configs/_base_/models/polyphonic_former.py:111-164."""
    s = {}

    def ln(p):
        s[p + '.weight'] = (C,)
        s[p + '.bias'] = (C,)

    def lin(p, o, i, bias=True):
        s[p + '.weight'] = (o, i)
        if bias:
            s[p + '.bias'] = (o,)

    for a in ('attention', 'attention_depth'):
        s[a + '.attn.in_proj_weight'] = (3 * C, C)
        s[a + '.attn.in_proj_bias'] = (3 * C,)
        lin(a + '.attn.out_proj', C, C)
    ln('attention_norm')
    ln('attention_norm_depth')
    for u in ('kernel_update_conv', 'kernel_update_conv_depth'):
        lin(u + '.dynamic_layer', 2 * C, C)
        lin(u + '.input_layer', 2 * C, C)
        lin(u + '.input_gate', C, C)
        lin(u + '.update_gate', C, C)
        for n in ('norm_in', 'norm_out', 'input_norm_in', 'input_norm_out'):
            ln(u + '.' + n)
        lin(u + '.fc_layer', C, C)
        ln(u + '.fc_norm')
    for t in ('feat_transform', 'feat_depth_transform'):
        s[t + '.conv.weight'] = (C, C, 1, 1)
        s[t + '.conv.bias'] = (C,)
    for f, n in (('ffn', 'ffn_norm'), ('ffn_depth', 'ffn_norm_depth')):
        lin(f + '.layers.0.0', FFN, C)
        lin(f + '.layers.1', C, FFN)
        ln(n)
    lin('cls_fcs.0', C, C, bias=False)
    ln('cls_fcs.1')
    lin('fc_cls', NUM_CLASSES, C)
    lin('mask_fcs.0', C, C, bias=False)
    ln('mask_fcs.1')
    lin('depth_regs.0', C, C, bias=False)
    ln('depth_regs.1')
    lin('fc_mask', C, C)
    lin('fc_depth', C, C)
    return s


def synth_stage_state(stage, seed=0):
    return {k: synth_tensor(f'stage{stage}.{k}', shp, seed)
            for k, shp in stage_state_shapes().items()}


def synth_decoder_state(num_stages=3, seed=0, prefix='mask_head.'):
    """state_dict of a KernelUpdateIterHead (keys ``mask_head.{s}.<...>``)."""
    out = {}
    for s in range(num_stages):
        for k, v in synth_stage_state(s, seed).items():
            out[f'{prefix}{s}.{k}'] = v
    return out


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def synth_decoder_inputs(B, H, W, seed=0, n_kernels=N_KERNELS):
    """Inputs of KernelUpdateIterHead.simple_test as KernelHead would hand them
    over (kernel_head.py:347): feature maps are sums of two post-ReLU maps, mask
    logits have ~40% of bits on.  Feature maps are pre-rounded to bf16 so the
    reference (fp32) and the bf16 device path see identical values."""
    g = _gen(f'inputs.{B}.{H}.{W}', seed)
    r = lambda *shape: torch.randn(*shape, generator=g)
    x = bf16_round(torch.relu(r(B, C, H, W)) + torch.relu(r(B, C, H, W)))
    d = bf16_round(torch.relu(r(B, C, H, W)) + 0.5 * torch.relu(r(B, C, H, W)))
    # low-frequency-ish mask logits: coarse noise upsampled + fine noise, shifted negative
    coarse = r(B, n_kernels, max(H // 4, 1), max(W // 4, 1))
    mask = torch.nn.functional.interpolate(coarse, size=(H, W), mode='bilinear',
                                           align_corners=False) * 3 + r(B, n_kernels, H, W) - 0.4
    prop = r(B, n_kernels, C, 1, 1) * 0.5
    dprop = (r(1, 1, C, 1, 1) * 0.1).expand(B, n_kernels, C, 1, 1).contiguous()
    dpred = r(B, 1, H, W)
    return dict(x_feats=x, depth_feats=d, mask_preds=mask.contiguous(), proposal_feats=prop,
                depth_proposal=dprop, depth_pred=dpred)


def synth_panoptic_inputs(h, w, seed=0, n_kernels=N_KERNELS):
    """Hand-constructed inputs of ``KernelUpdateIterHead.get_panoptic`` (kernel_update.py:421-469): with random
    weights the decoder's own outputs merge into ~1 segment, so parity of the post-processing is pinned on these
    instead.  Thing masks are blobs, stuff masks are bands; a dozen things and most stuff classes score high; no two
    scores are equal.  Returns cls_scores [N,19] (already sigmoid), mask / depth logits [N,h,w], depth_init [1,h,w]."""
    g = _gen(f'panoptic.{h}.{w}', seed)
    r = lambda *shape: torch.rand(*shape, generator=g)
    ys = torch.arange(h, dtype=torch.float32).view(1, h, 1)
    xs = torch.arange(w, dtype=torch.float32).view(1, 1, w)
    cy, cx = r(N_PROPOSALS, 1, 1) * h, r(N_PROPOSALS, 1, 1) * w
    rad = 2.0 + r(N_PROPOSALS, 1, 1) * (h / 4)
    thing = 5.0 - 7.0 * ((ys - cy) ** 2 + (xs - cx) ** 2) / rad ** 2
    band0 = torch.linspace(0, h, NUM_STUFF + 1)
    stuff = torch.stack([2.5 - 2.0 * ((ys[0] - (band0[i] + band0[i + 1]) / 2).abs() / (h / NUM_STUFF)).expand(h, w)
                         for i in range(NUM_STUFF)])
    mask = torch.cat([thing, stuff])[:n_kernels] + 0.05 * torch.randn(n_kernels, h, w, generator=g)
    cls = 0.01 + 0.08 * r(n_kernels, NUM_CLASSES)
    strong = torch.randperm(N_PROPOSALS, generator=g)[:14]
    for j, n in enumerate(strong.tolist()):
        cls[n, j % NUM_THING] = 0.35 + 0.6 * r(1).item() if j != 3 else 0.2      # one below instance_score_thr
    for i in range(NUM_STUFF):
        cls[N_PROPOSALS + i, NUM_THING + i] = 0.15 + 0.8 * r(1).item()
    coarse = torch.randn(n_kernels, max(h // 8, 1), max(w // 8, 1), generator=g)
    depth = torch.nn.functional.interpolate(coarse[None], size=(h, w), mode='bilinear', align_corners=False)[0]
    dinit = torch.nn.functional.interpolate(torch.randn(1, 1, max(h // 8, 1), max(w // 8, 1), generator=g), size=(h, w),
                                            mode='bilinear', align_corners=False)[0]
    return dict(cls_scores=cls, mask_preds=mask.contiguous(), depth_preds=depth.contiguous(), depth_init=dinit.contiguous())


def kernel_head_state_shapes(num_proposals=N_PROPOSALS):
    """name -> shape of the KernelHead parameters AFTER SemanticFPN (kernel_head.py:142-200)."""
    s = {}
    for m in ('loc', 'seg', 'depth'):
        s[f'{m}_convs.0.conv.weight'] = (C, C, 1, 1)
        s[f'{m}_convs.0.gn.weight'] = (C,)
        s[f'{m}_convs.0.gn.bias'] = (C,)
    s['init_kernels.weight'] = (num_proposals, C, 1, 1)
    s['conv_seg.weight'] = (NUM_CLASSES, C, 1, 1)
    s['conv_seg.bias'] = (NUM_CLASSES,)
    s['conv_direct_depth.weight'] = (1, C, 1, 1)
    s['conv_direct_depth.bias'] = (1,)
    return s


def synth_kernel_head_state(seed=0, num_proposals=N_PROPOSALS):
    return {k: synth_tensor('kernel_head.' + k, shp, seed) for k, shp in kernel_head_state_shapes(num_proposals).items()}


def synth_fpn_maps(B, H, W, seed=0):
    """The three SemanticFPN outputs (localization, semantic, depth; each the ReLU output of a conv+GN stack,
    semantic_fpn.py:221-229), pre-rounded to bf16 -- the storage dtype both sides consume."""
    g = _gen(f'fpn.{B}.{H}.{W}', seed)
    maps = []
    for scale in (1.0, 0.8, 1.3):
        coarse = torch.randn(B, C, max(H // 4, 1), max(W // 4, 1), generator=g)
        t = torch.nn.functional.interpolate(coarse, size=(H, W), mode='bilinear', align_corners=False)
        maps.append(bf16_round(torch.relu(scale * (t + 0.7 * torch.randn(B, C, H, W, generator=g)) + 0.2)))
    return maps


def synth_fpn_pred_state(seed=0):
    """conv_pred + two aux_convs of SemanticFPNWrapper (semantic_fpn.py:159-178)."""
    out = {}
    for n in ('conv_pred', 'aux_convs.0', 'aux_convs.1'):
        out[n + '.conv.weight'] = synth_tensor('fpn.' + n + '.conv.weight', (C, C, 1, 1), seed)
        out[n + '.gn.weight'] = synth_tensor('fpn.' + n + '.gn.weight', (C,), seed)
        out[n + '.gn.bias'] = synth_tensor('fpn.' + n + '.gn.bias', (C,), seed)
    return out


def synth_fused_map(B, H, W, seed=0):
    """feature_add_all_level: the sum of four post-ReLU level maps (semantic_fpn.py:216-219), pre-rounded to bf16."""
    g = _gen(f'fused.{B}.{H}.{W}', seed)
    t = sum(torch.relu(torch.randn(B, C, H, W, generator=g) * s) for s in (1.0, 0.7, 0.5, 0.4))
    return bf16_round(t)


def synth_semantic_fpn_state(seed=0):
    """All 30 tensors of SemanticFPNWrapper (semantic_fpn.py:74-178 with the shipped config)."""
    out = synth_fpn_pred_state(seed)
    for lvl, n in ((0, 1), (1, 1), (2, 2), (3, 3)):
        for j in range(n):
            p = f'convs_all_levels.{lvl}.conv{j}'
            out[p + '.conv.weight'] = synth_tensor('fpn.' + p + '.conv.weight', (C, C, 3, 3), seed)
            out[p + '.gn.weight'] = synth_tensor('fpn.' + p + '.gn.weight', (C,), seed)
            out[p + '.gn.bias'] = synth_tensor('fpn.' + p + '.gn.bias', (C,), seed)
    return out


def synth_fpn_inputs(B, H, W, seed=0):
    """The four FPN levels (strides 4, 8, 16, 32 of the frame) for a decoder map of H x W (stride 8)."""
    g = _gen(f'fpnin.{B}.{H}.{W}', seed)
    return [torch.randn(B, C, h, w, generator=g) for h, w in ((2 * H, 2 * W), (H, W), (H // 2, W // 2), (H // 4, W // 4))]


def synth_track_head_state(seed=0):
    """QuasiDenseMaskEmbedHeadGTMask (track_heads.py:14-76 with configs/polyphonic_video/poly_r50_cityscapes_1x.py:38-52):
    4 x (3x3 conv + GN), FC 12544 -> 1024, FC 1024 -> 256.  fc_embed is scaled up so that the bi-softmax association of
    the tracker is decisive on random weights."""
    out = {}
    for i in range(4):
        out[f'convs.{i}.conv.weight'] = synth_tensor(f'track.convs.{i}.conv.weight', (C, C, 3, 3), seed)
        out[f'convs.{i}.gn.weight'] = synth_tensor(f'track.convs.{i}.gn.weight', (C,), seed)
        out[f'convs.{i}.gn.bias'] = synth_tensor(f'track.convs.{i}.gn.bias', (C,), seed)
    out['fcs.0.weight'] = synth_tensor('track.fcs.0.weight', (1024, C * 49), seed)
    out['fcs.0.bias'] = synth_tensor('track.fcs.0.bias', (1024,), seed)
    out['fc_embed.weight'] = synth_tensor('track.fc_embed.weight', (C, 1024), seed) * 4.0
    out['fc_embed.bias'] = synth_tensor('track.fc_embed.bias', (C,), seed)
    return out


def synth_clip(n_frames=4, H=128, W=192, seed=0):
    """A synthetic clip for the tracking path: per frame the four FPN levels (a static texture plus a little noise, so
    that the same object embeds alike in neighbouring frames) and K thing masks [K,H,W] drifting across the frame,
    with class labels and scores.  Object 1 overlaps object 0 almost completely (duplicate removal), object 3 has a low
    score, object 4 appears in frame 1 only from the second frame on, object 5 is always empty."""
    g = _gen(f'clip.{n_frames}.{H}.{W}', seed)
    base = [torch.randn(1, C, H // s, W // s, generator=g) for s in (4, 8, 16, 32)]
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing='ij')
    frames = []
    for t in range(n_frames):
        feats = [b + 0.05 * torch.randn(b.shape, generator=g) for b in base]
        objs = [  # cy, cx, ry, rx, label, score
            (30 + 3 * t, 40 + 5 * t, 14, 22, 2, 0.92), (31 + 3 * t, 42 + 5 * t, 13, 21, 2, 0.71),
            (90 - 2 * t, 120 + 4 * t, 20, 16, 5, 0.64), (60, 150 - 6 * t, 9, 9, 5, 0.22),
            (100, 30 + 8 * t, 12, 18, 0, 0.55), (0, 0, 0, 0, 1, 0.40)]
        masks, labels, scores = [], [], []
        for k, (cy, cx, ry, rx, lab, sc) in enumerate(objs):
            if k == 4 and t == 0:
                continue
            m = ((yy - cy).float() / max(ry, 1)) ** 2 + ((xx - cx).float() / max(rx, 1)) ** 2 <= 1.0 if ry else torch.zeros(H, W, dtype=torch.bool)
            masks.append(m), labels.append(lab), scores.append(sc)
        frames.append(dict(feats=feats, masks=torch.stack(masks), labels=torch.tensor(labels), scores=torch.tensor(scores)))
    return frames
