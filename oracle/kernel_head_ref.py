"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch fp32) of the tail of
``KernelHead._decode_init_proposals`` (polyphonic/kernel_head.py:250-336 of the reference), from the three
SemanticFPN maps on.  Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this module.

Pinned against the real reference by tests/test_oracle_golden.py (fixtures from oracle/make_golden.py, which runs
the reference's own KernelHead with ``localization_fpn`` replaced by the synthetic maps).
"""
import torch
import torch.nn.functional as F


def conv_gn_relu(sd, name, x, num_groups=32, eps=1e-5):
    """mmcv ConvModule(256, 256, 1, norm_cfg=GN32): conv without bias -> GroupNorm -> ReLU (kernel_head.py:173-198)."""
    y = F.conv2d(x, sd[f'{name}.0.conv.weight'])
    y = F.group_norm(y, num_groups, sd[f'{name}.0.gn.weight'], sd[f'{name}.0.gn.bias'], eps)
    return F.relu(y)


def fpn_pred(sd, fused, num_groups=32, eps=1e-5):
    """SemanticFPNWrapper.forward, polyphonic/funcs/semantic_fpn.py:221-229 with num_aux_convs=2: conv_pred and the two
    aux_convs (ConvModule 1x1 + GN32 + ReLU, built :159-178) on the fused map.  Returns [out, aux0, aux1]."""
    outs = []
    for name in ('conv_pred', 'aux_convs.0', 'aux_convs.1'):
        y = F.conv2d(fused, sd[name + '.conv.weight'])
        y = F.group_norm(y, num_groups, sd[name + '.gn.weight'], sd[name + '.gn.bias'], eps)
        outs.append(F.relu(y))
    return outs


def decode_init_proposals(sd, localization_feats, num_thing_classes=8):
    """kernel_head.py:250-336 in eval mode with the shipped configuration (cat_stuff_mask, use_binary,
    proposal_feats_with_obj, no feat_refine, no semantic_aspp).  Returns the reference's 9-tuple as a dict."""
    loc, sem, dep = localization_feats
    B = loc.shape[0]
    loc_feats = conv_gn_relu(sd, 'loc_convs', loc)                                   # :250-251
    mask_preds = F.conv2d(loc_feats, sd['init_kernels.weight'])                      # :256
    semantic_feats = conv_gn_relu(sd, 'seg_convs', sem)                              # :264-265
    depth_feats = conv_gn_relu(sd, 'depth_convs', dep)                               # :277-278
    depth_pred = F.conv2d(depth_feats, sd['conv_direct_depth.weight'], sd['conv_direct_depth.bias'])   # :285
    seg_preds = F.conv2d(semantic_feats, sd['conv_seg.weight'], sd['conv_seg.bias'])                   # :295
    x_feats = semantic_feats + loc_feats                                             # :303
    binary = (mask_preds.sigmoid() > 0.5).float()                                    # :314-317
    obj_feats = torch.einsum('bnhw,bchw->bnc', binary, x_feats)                      # :320
    P, C = sd['init_kernels.weight'].shape[:2]
    proposal_feats = sd['init_kernels.weight'][None].expand(B, P, C, 1, 1) + obj_feats.view(B, P, C, 1, 1)   # :299-326
    T = num_thing_classes
    mask_preds = torch.cat([mask_preds, seg_preds[:, T:]], dim=1)                    # :329-331
    stuff_kernels = sd['conv_seg.weight'][T:][None].expand(B, -1, C, 1, 1)           # :332-334
    proposal_feats = torch.cat([proposal_feats, stuff_kernels], dim=1)               # :335
    N = proposal_feats.shape[1]
    depth_proposal = sd['conv_direct_depth.weight'][None].expand(B, 1, C, 1, 1).expand(-1, N, -1, -1, -1)   # :286-289, :336
    return dict(proposal_feats=proposal_feats, x_feats=x_feats, mask_preds=mask_preds, cls_scores=None,
                seg_preds=seg_preds, depth_feats=depth_feats, depth_proposal=depth_proposal, depth_pred=depth_pred,
                semantic_aspp_out=None)
