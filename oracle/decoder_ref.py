"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch fp32) of the
PolyphonicFormer decoder hot path.  The product package never imports this file;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs do, and there only as the checker / the timed baseline.

PARITY PIN: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so this restatement is pinned against outputs of the
reference's own unmodified code run in the build container through
``oracle/mmcv_shim.py`` -- see ``oracle/make_golden.py`` and
``tests/test_oracle_golden.py`` (fixtures in ``tests/golden/``).

Every function cites the reference lines (relative to /root/reference) it follows.
All functions are functional: parameters come from a ``state_dict``-style mapping
with the reference's own key names.
"""
import math

import torch
import torch.nn.functional as F

LN_EPS = 1e-5   # mmcv build_norm_layer(dict(type='LN')) -> nn.LayerNorm default eps


def _sub(sd, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def _ln(x, sd, name):
    return F.layer_norm(x, (x.shape[-1],), sd[name + '.weight'], sd[name + '.bias'], LN_EPS)


def _lin(x, sd, name):
    return F.linear(x, sd[name + '.weight'], sd.get(name + '.bias'))


def kernel_updator(sd, update_feature, input_feature):
    """polyphonic/funcs/kernel_updator.py:55-93 with the shipped config
    (feat_channels = in_channels = out_channels = 256, gate_sigmoid=True,
    gate_norm_act=False, activate_out=False; configs/_base_/models/polyphonic_former.py:131-138).

    update_feature: [R, 256] pooled feature, input_feature: [R, 1, 256] kernel.
    Returns [R, 1, 256].
    """
    Cc = sd['fc_layer.weight'].shape[1]
    update_feature = update_feature.reshape(-1, Cc)                        # :56
    R = update_feature.size(0)
    parameters = _lin(update_feature, sd, 'dynamic_layer')                 # :58
    param_in = parameters[:, :Cc]                                          # :59-60
    param_out = parameters[:, -Cc:]                                        # :61-62
    input_feats = _lin(input_feature.reshape(R, -1, Cc), sd, 'input_layer')  # :64-65
    input_in = input_feats[..., :Cc]                                       # :66
    input_out = input_feats[..., -Cc:]                                     # :67
    gate_feats = input_in * param_in.unsqueeze(-2)                         # :69
    input_gate = _ln(_lin(gate_feats, sd, 'input_gate'), sd, 'input_norm_in').sigmoid()   # :73,75-76
    update_gate = _ln(_lin(gate_feats, sd, 'update_gate'), sd, 'norm_in').sigmoid()       # :74,77
    param_out = _ln(param_out, sd, 'norm_out')                             # :78
    input_out = _ln(input_out, sd, 'input_norm_out')                       # :79
    features = update_gate * param_out.unsqueeze(-2) + input_gate * input_out   # :86-87
    features = _lin(features, sd, 'fc_layer')                              # :89
    features = _ln(features, sd, 'fc_norm')                                # :90
    return torch.relu(features)                                            # :91


def multihead_self_attention(sd, x, num_heads=8):
    """mmcv 1.3.18 ``MultiheadAttention`` wrapper around ``nn.MultiheadAttention``
    as called at polyphonic/kernel_update_head.py:259-260: sequence-first input
    [L, B, E], q = k = v = x, returns ``x + out_proj(softmax(q k^T / sqrt(d)) v)``.
    Restated from torch ``F.multi_head_attention_forward`` (torch/nn/functional.py):
    packed in-projection, q scaled by 1/sqrt(d) before the product, softmax over keys.
    """
    L, B, E = x.shape
    d = E // num_heads
    qkv = F.linear(x, sd['attn.in_proj_weight'], sd['attn.in_proj_bias'])  # [L,B,3E]
    q, k, v = qkv.split(E, dim=-1)
    # [L, B, E] -> [B*heads, L, d]
    q = q.reshape(L, B * num_heads, d).transpose(0, 1) * (1.0 / math.sqrt(d))
    k = k.reshape(L, B * num_heads, d).transpose(0, 1)
    v = v.reshape(L, B * num_heads, d).transpose(0, 1)
    attn = torch.softmax(torch.bmm(q, k.transpose(1, 2)), dim=-1)
    out = torch.bmm(attn, v).transpose(0, 1).reshape(L, B, E)
    out = F.linear(out, sd['attn.out_proj.weight'], sd['attn.out_proj.bias'])
    return x + out                                                         # mmcv: identity + attn


def ffn(sd, x):
    """mmcv 1.3.18 ``FFN`` (num_fcs=2, ReLU, dropout 0, add_identity) as called at
    polyphonic/kernel_update_head.py:271-272: ``x + W2 relu(W1 x + b1) + b2``."""
    h = torch.relu(_lin(x, sd, 'layers.0.0'))
    return x + _lin(h, sd, 'layers.1')


def binarise(mask_logits, hard_mask_thr=0.5):
    """polyphonic/kernel_update_head.py:236-238."""
    return (mask_logits.sigmoid() > hard_mask_thr).float()


def kernel_update_head(sd, x, proposal_feat, mask_preds, depth_proposal, depth_feats,
                       num_heads=8):
    """One decoder stage: polyphonic/kernel_update_head.py:212-353 with the shipped
    config (conv_kernel_size=1, feat_gather_stride=1, mask_transform_stride=1,
    with_ffn, num_cls_fcs=num_mask_fcs=1; 1x1 ``feat_transform`` with bias and no
    norm/activation, :124-140).

    x, depth_feats: [B,C,H,W]; proposal_feat, depth_proposal: [B,N,C,1,1];
    mask_preds: [B,N,H,W] logits.
    Returns (cls_score [B,N,classes], new_mask_preds [B,N,H,W], obj_feat [B,N,C,1,1],
             new_depth_preds [B,N,H,W], depth_feat_new [B,N,C,1,1]).
    """
    B, N = proposal_feat.shape[:2]
    x = F.conv2d(x, sd['feat_transform.conv.weight'], sd['feat_transform.conv.bias'])          # :225
    depth_feats = F.conv2d(depth_feats, sd['feat_depth_transform.conv.weight'],
                           sd['feat_depth_transform.conv.bias'])                                # :226
    C, H, W = x.shape[-3:]
    if mask_preds.shape[-2:] != (H, W):                                                         # :229-234
        mask_preds = F.interpolate(mask_preds, (H, W), align_corners=False, mode='bilinear')
    m = binarise(mask_preds)                                                                    # :236-238
    x_feat = torch.einsum('bnhw,bchw->bnc', m, x)                                               # :241
    d_feat = torch.einsum('bnhw,bchw->bnc', m, depth_feats)                                     # :242

    proposal_feat = proposal_feat.reshape(B, N, C, -1).permute(0, 1, 3, 2)                      # :245-247
    depth_proposal = depth_proposal.reshape(B, N, C, -1).permute(0, 1, 3, 2)                    # :248-249
    depth_proposal = depth_proposal + proposal_feat                                             # :250

    obj = kernel_updator(_sub(sd, 'kernel_update_conv.'), x_feat, proposal_feat)                # :252
    dep = kernel_updator(_sub(sd, 'kernel_update_conv_depth.'), d_feat, depth_proposal)         # :253

    obj = obj.reshape(B, N, -1).permute(1, 0, 2)                                                # :256
    dep = dep.reshape(B, N, -1).permute(1, 0, 2)                                                # :257
    obj = _ln(multihead_self_attention(_sub(sd, 'attention.'), obj, num_heads), sd, 'attention_norm')            # :259
    dep = _ln(multihead_self_attention(_sub(sd, 'attention_depth.'), dep, num_heads), sd, 'attention_norm_depth')  # :260
    obj = obj.permute(1, 0, 2).reshape(B, N, -1, C)                                             # :262,266
    dep = dep.permute(1, 0, 2).reshape(B, N, -1, C)                                             # :263,267

    obj = _ln(ffn(_sub(sd, 'ffn.'), obj), sd, 'ffn_norm')                                       # :271
    dep = _ln(ffn(_sub(sd, 'ffn_depth.'), dep), sd, 'ffn_norm_depth')                           # :272

    cls_feat = obj.sum(-2)                                                                      # :274
    cls_feat = torch.relu(_ln(F.linear(cls_feat, sd['cls_fcs.0.weight']), sd, 'cls_fcs.1'))     # :278-279
    mask_feat = torch.relu(_ln(F.linear(obj, sd['mask_fcs.0.weight']), sd, 'mask_fcs.1'))       # :280-281
    depth_k = _ln(F.linear(dep, sd['depth_regs.0.weight']), sd, 'depth_regs.1')                 # :282-283 (no act)

    cls_score = _lin(cls_feat, sd, 'fc_cls').view(B, N, -1)                                     # :285
    mask_feat = _lin(mask_feat, sd, 'fc_mask').reshape(B, N, C)                                 # :287,308
    depth_k = _lin(depth_k, sd, 'fc_depth').reshape(B, N, C)                                    # :288,311

    # per-image 1x1 conv with dynamic weights == batched matmul                                 # :317-334
    new_mask = torch.einsum('bnc,bchw->bnhw', mask_feat, x)
    new_depth = torch.einsum('bnc,bchw->bnhw', depth_k, depth_feats)

    obj_out = obj.permute(0, 1, 3, 2).reshape(B, N, C, 1, 1)                                    # :349-350
    dep_out = dep.permute(0, 1, 3, 2).reshape(B, N, C, 1, 1)                                    # :351-353
    return cls_score, new_mask, obj_out, new_depth, dep_out


def upsample2x(t):
    """F.interpolate(scale_factor=2, bilinear, align_corners=False),
    polyphonic/kernel_update.py:133-143."""
    return F.interpolate(t, scale_factor=2, mode='bilinear', align_corners=False)


def decoder_forward(sd, x_feats, proposal_feats, mask_preds, depth_feats, depth_proposal,
                    num_stages=3, mask_upsample_stride=2, prefix='mask_head.',
                    return_all_stages=False):
    """Stage loop of KernelUpdateIterHead.simple_test up to (not including) the
    post-processing: polyphonic/kernel_update.py:316-336 and _mask_forward :125-157.

    Returns dict with cls_score (sigmoid applied, :333-334), mask_preds,
    scaled_mask_preds, depth_preds, scaled_depth_preds, object_feats, depth_proposal.
    """
    object_feats = proposal_feats
    stages = []
    for s in range(num_stages):
        ssd = _sub(sd, f'{prefix}{s}.')
        cls_score, mask_preds, object_feats, depth_preds, depth_proposal = kernel_update_head(
            ssd, x_feats, object_feats, mask_preds, depth_proposal, depth_feats)
        if return_all_stages:
            stages.append(dict(cls_score=cls_score, mask_preds=mask_preds, object_feats=object_feats,
                               depth_preds=depth_preds, depth_proposal=depth_proposal))
    if mask_upsample_stride > 1:                                            # kernel_update.py:131-143
        scaled_mask = upsample2x(mask_preds)
        scaled_depth = upsample2x(depth_preds)
    else:
        scaled_mask, scaled_depth = mask_preds, depth_preds
    out = dict(cls_score=cls_score.sigmoid(), mask_preds=mask_preds, scaled_mask_preds=scaled_mask,
               depth_preds=depth_preds, scaled_depth_preds=scaled_depth,
               object_feats=object_feats, depth_proposal=depth_proposal)
    if return_all_stages:
        out['stages'] = stages
    return out
