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
