"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch fp32) of ``SemanticFPNWrapper.forward``
(polyphonic/funcs/semantic_fpn.py:198-235 of the reference) in the shipped configuration
(configs/_base_/models/polyphonic_former.py:75-95: levels 0..3, upsample_times=2, sine positional encoding added at
level 3, no coordinate channels, sum fusion, conv_pred + two aux_convs).  SURVEY.md section 8f rank 4: the oracle for
the kernels that will replace the 3x3 conv + GroupNorm + bilinear pyramid; pinned against the real module by
tests/test_oracle_golden.py (fixture from oracle/make_golden.py).  Only tests/, smoke() and bench.py's cpu legs may
import this module.
"""
import math

import torch
import torch.nn.functional as F


def sine_positional_encoding(B, H, W, num_feats=128, temperature=10000, scale=2 * math.pi, eps=1e-6):
    """mmdet/models/utils/positional_encoding.py:57-92 with normalize=True, offset=0 and an all-valid mask."""
    ones = torch.ones(B, H, W, dtype=torch.float32)
    y_embed, x_embed = ones.cumsum(1), ones.cumsum(2)
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / num_feats)
    pos_x, pos_y = x_embed[..., None] / dim_t, y_embed[..., None] / dim_t
    pos_x = torch.stack((pos_x[..., 0::2].sin(), pos_x[..., 1::2].cos()), dim=4).view(B, H, W, -1)
    pos_y = torch.stack((pos_y[..., 0::2].sin(), pos_y[..., 1::2].cos()), dim=4).view(B, H, W, -1)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


def conv3x3_gn_relu(sd, name, x, stride=1, num_groups=32, eps=1e-5):
    """mmcv ConvModule(256, 256, 3, padding=1, norm_cfg=GN32, act ReLU): conv without bias -> GroupNorm -> ReLU."""
    y = F.conv2d(x, sd[name + '.conv.weight'], stride=stride, padding=1)
    return F.relu(F.group_norm(y, num_groups, sd[name + '.gn.weight'], sd[name + '.gn.bias'], eps))


def up2(x):
    return F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)   # nn.Upsample, semantic_fpn.py:125-129


def fused_levels(sd, inputs):
    """semantic_fpn.py:199-219: per-level stacks (built :74-152) and their sum.  inputs: 4 FPN maps, strides 4/8/16/32."""
    p0, p1, p2, p3 = inputs
    l0 = conv3x3_gn_relu(sd, 'convs_all_levels.0.conv0', p0, stride=2)            # level 0: one stride-2 conv (:92-104)
    l1 = conv3x3_gn_relu(sd, 'convs_all_levels.1.conv0', p1)                      # level 1: one conv, no upsample
    l2 = conv3x3_gn_relu(sd, 'convs_all_levels.2.conv0', p2)                      # level 2: conv, x2, conv
    l2 = conv3x3_gn_relu(sd, 'convs_all_levels.2.conv1', up2(l2))
    B, _, H, W = p3.shape
    l3 = p3 + sine_positional_encoding(B, H, W)                                    # :203-210 (cat_coors_level = 3)
    l3 = conv3x3_gn_relu(sd, 'convs_all_levels.3.conv0', l3)                      # level 3: conv, x2, conv, x2, conv
    l3 = conv3x3_gn_relu(sd, 'convs_all_levels.3.conv1', up2(l3))
    l3 = conv3x3_gn_relu(sd, 'convs_all_levels.3.conv2', up2(l3))
    return l0 + l1 + l2 + l3                                                       # :216-219 (fuse_by_cat=False)


def semantic_fpn_forward(sd, inputs):
    """The whole forward: fused map -> [conv_pred, aux_convs.0, aux_convs.1] (:221-229, kernel_head_ref.fpn_pred)."""
    from oracle.kernel_head_ref import fpn_pred
    fused = fused_levels(sd, inputs)
    return fused, fpn_pred(sd, fused)
