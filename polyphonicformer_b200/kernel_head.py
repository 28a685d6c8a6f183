"""Host side of K5: the tail of ``KernelHead._decode_init_proposals`` (polyphonic/kernel_head.py:250-336 of the
reference) -- everything between SemanticFPN and the decoder's stage loop -- on the CUDA kernels behind
``pf_kernel_head`` / ``pf_mask_pool`` / ``pf_init_proposals`` (include/pf_decoder.h).  PyTorch is device memory and
streams only; there is no fallback.

State-dict keys consumed (prefix ``rpn_head.`` in a full model):
  {loc,seg,depth}_convs.0.conv.weight [256,256,1,1], {loc,seg,depth}_convs.0.gn.{weight,bias} [256],
  init_kernels.weight [P,256,1,1], conv_seg.{weight [19,256,1,1], bias}, conv_direct_depth.{weight [1,256,1,1], bias}
"""
import ctypes

import torch

from . import _cabi
from ._cabi import PF_C, FpnWeights, HeadWeights
from .decoder import _ptr, _stream_ptr, round_up

H_ROWS, H_ROW_SEG, H_ROW_DEP = 160, 112, 144   # head_w row blocks (include/pf_decoder.h)


def _hi_lo(w32):
    hi = w32.to(torch.bfloat16)
    lo = (w32 - hi.to(torch.float32)).to(torch.bfloat16)
    return hi, lo


class PackedKernelHead:
    """The static weights of the KernelHead tail in the layouts ``struct pf_head_weights`` documents."""

    def __init__(self, sd, device, num_thing_classes=8, gn_eps=1e-5):
        g = lambda k: sd[k].detach().to('cpu', torch.float32)
        conv = [g(f'{m}_convs.0.conv.weight').reshape(PF_C, PF_C) for m in ('loc', 'seg', 'depth')]
        self.conv_split = _conv_split(conv).to(device)                            # [6*2][128][256]
        self.gn_gamma = torch.stack([g(f'{m}_convs.0.gn.weight') for m in ('loc', 'seg', 'depth')]).contiguous().to(device)
        self.gn_beta = torch.stack([g(f'{m}_convs.0.gn.bias') for m in ('loc', 'seg', 'depth')]).contiguous().to(device)
        init_k = g('init_kernels.weight').reshape(-1, PF_C)
        seg_w = g('conv_seg.weight').reshape(-1, PF_C)
        dep_w = g('conv_direct_depth.weight').reshape(1, PF_C)
        self.num_proposals, self.num_classes = init_k.shape[0], seg_w.shape[0]
        self.num_thing_classes = num_thing_classes
        if self.num_proposals > H_ROW_SEG or self.num_classes > H_ROW_DEP - H_ROW_SEG:
            raise ValueError('pf_kernel_head supports at most %d proposals and %d classes' % (H_ROW_SEG, H_ROW_DEP - H_ROW_SEG))
        w = torch.zeros((H_ROWS, PF_C), dtype=torch.float32)
        w[:self.num_proposals] = init_k
        w[H_ROW_SEG:H_ROW_SEG + self.num_classes] = seg_w
        w[H_ROW_DEP] = dep_w[0]
        hi, lo = _hi_lo(w)
        self.head_w = torch.stack([hi, lo]).contiguous().to(device)               # [2][160][256]
        b = torch.zeros(H_ROWS, dtype=torch.float32)
        b[H_ROW_SEG:H_ROW_SEG + self.num_classes] = g('conv_seg.bias')
        b[H_ROW_DEP] = g('conv_direct_depth.bias')[0]
        self.head_b = b.to(device)
        # fp32 operands of pf_init_proposals and of the depth-kernel expansion (kernel_head.py:286-289, 329-336)
        self.init_kernels = init_k.contiguous().to(device)
        self.stuff_kernels = seg_w[num_thing_classes:].contiguous().to(device)
        self.depth_kernel = dep_w.contiguous().to(device)
        self.struct = HeadWeights(conv_split=self.conv_split.data_ptr(), gn_gamma=self.gn_gamma.data_ptr(),
                                  gn_beta=self.gn_beta.data_ptr(), head_w=self.head_w.data_ptr(),
                                  head_b=self.head_b.data_ptr(), num_proposals=self.num_proposals,
                                  num_classes=self.num_classes, num_thing_classes=num_thing_classes, gn_eps=gn_eps)


def _conv_split(convs):
    """three fp32 [256][256] 1x1-conv matrices -> bf16 [6*2][128][256]: block (half * 3 + map), hi plane then lo plane"""
    blocks = []
    for half in range(2):
        for m in range(3):
            hi, lo = _hi_lo(convs[m][128 * half:128 * half + 128])
            blocks += [hi, lo]
    return torch.stack(blocks).contiguous()


class FpnPred:
    """The last step of ``SemanticFPNWrapper.forward`` (polyphonic/funcs/semantic_fpn.py:221-229): ``conv_pred`` and the
    two ``aux_convs`` (1x1 conv + GN32 + ReLU each) on the fused multi-level map -> the three ``localization_feats`` in
    the bf16 layout ``KernelHeadTail.forward`` consumes.  State-dict keys (prefix ``rpn_head.localization_fpn.``):
    conv_pred.{conv.weight, gn.weight, gn.bias}, aux_convs.{0,1}.{conv.weight, gn.weight, gn.bias}."""
    NAMES = ('conv_pred', 'aux_convs.0', 'aux_convs.1')

    def __init__(self, state_dict, device, gn_eps=1e-5):
        _cabi.load()
        self.device = torch.device(device)
        g = lambda k: state_dict[k].detach().to('cpu', torch.float32)
        self.conv_split = _conv_split([g(n + '.conv.weight').reshape(PF_C, PF_C) for n in self.NAMES]).to(self.device)
        self.gn_gamma = torch.stack([g(n + '.gn.weight') for n in self.NAMES]).contiguous().to(self.device)
        self.gn_beta = torch.stack([g(n + '.gn.bias') for n in self.NAMES]).contiguous().to(self.device)
        self.gn_eps = gn_eps
        self._ws = None

    def forward(self, fused, want_fp32=False, hw=None):
        """fused: fp32 [B,256,H,W] (feature_add_all_level), or -- with hw = (H, W) -- the bf16 [B][256][HWp] buffer
        SemanticFpnPyramid wrote.  Returns (maps bf16 [3][B][256][HWp], maps32 or None)."""
        lib = _cabi.load()
        st = _stream_ptr()
        if hw is not None:
            (H, W), B, fb = hw, fused.shape[0], fused
            HW = H * W
            HWp = round_up(HW, 8)
            assert fb.dtype == torch.bfloat16 and tuple(fb.shape) == (B, PF_C, HWp) and fb.is_contiguous()
        else:
            B, _, H, W = fused.shape
            HW = H * W
            HWp = round_up(HW, 8)
            fused = fused.to(self.device, torch.float32).contiguous()
            fb = torch.empty((B, PF_C, HWp), dtype=torch.bfloat16, device=self.device)
            _cabi.call('pf_cast_maps', _ptr(fused), _ptr(fb), B * PF_C, HW, HWp, st)
        maps = torch.empty((3, B, PF_C, HWp), dtype=torch.bfloat16, device=self.device)
        maps32 = torch.empty((3, B, PF_C, H, W), dtype=torch.float32, device=self.device) if want_fp32 else None
        nbytes = lib.pf_kernel_head_workspace_bytes(B, HW)
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        _cabi.call('pf_fpn_pred', _ptr(self.conv_split), _ptr(self.gn_gamma), _ptr(self.gn_beta), self.gn_eps, _ptr(fb),
                   _ptr(maps), _ptr(maps32), _ptr(self._ws), nbytes, B, HW, HWp, st)
        self.last_launches = lib.pf_last_launch_count()
        return maps, maps32


class SemanticFpnPyramid:
    """``SemanticFPNWrapper.forward`` up to ``feature_add_all_level`` (polyphonic/funcs/semantic_fpn.py:198-219) in the shipped
    configuration, on ``pf_semantic_fpn``: seven 3x3 conv + GN32 + ReLU modules as shifted-row tensor-core GEMMs, the x2
    bilinear steps and the four-level sum.  State-dict keys (prefix ``rpn_head.localization_fpn.``):
    convs_all_levels.{0.conv0, 1.conv0, 2.conv0, 2.conv1, 3.conv0, 3.conv1, 3.conv2}.{conv.weight, gn.weight, gn.bias}."""
    NAMES = ('convs_all_levels.0.conv0', 'convs_all_levels.1.conv0', 'convs_all_levels.2.conv0', 'convs_all_levels.2.conv1',
             'convs_all_levels.3.conv0', 'convs_all_levels.3.conv1', 'convs_all_levels.3.conv2')

    def __init__(self, state_dict, device, gn_eps=1e-5):
        _cabi.load()
        self.device = torch.device(device)
        g = lambda k: state_dict[k].detach().to('cpu', torch.float32)
        planes = []
        for n in self.NAMES:
            w = g(n + '.conv.weight')
            if tuple(w.shape) != (PF_C, PF_C, 3, 3):
                raise NotImplementedError('%s.conv.weight has shape %s, expected 256x256x3x3' % (n, tuple(w.shape)))
            hi, lo = _hi_lo(w.permute(2, 3, 0, 1).reshape(9, PF_C, PF_C))          # [ky*3+kx][out][in]
            planes += [hi, lo]
        self.conv_w = torch.stack(planes).contiguous().to(self.device)            # [7*2][9][256][256]
        self.gn_gamma = torch.stack([g(n + '.gn.weight') for n in self.NAMES]).contiguous().to(self.device)
        self.gn_beta = torch.stack([g(n + '.gn.bias') for n in self.NAMES]).contiguous().to(self.device)
        self.struct = FpnWeights(conv_w=self.conv_w.data_ptr(), gn_gamma=self.gn_gamma.data_ptr(),
                                 gn_beta=self.gn_beta.data_ptr(), gn_eps=gn_eps)
        self._ws = None

    def forward(self, inputs, want_fp32=False):
        """inputs: the four FPN levels fp32 [B,256,2H,2W], [B,256,H,W], [B,256,H/2,W/2], [B,256,H/4,W/4].
        Returns (fused bf16 [B][256][HWp], fused fp32 [B,256,H,W] or None, (H, W))."""
        lib = _cabi.load()
        p = [t.to(self.device, torch.float32).contiguous() for t in inputs[:4]]
        B, C, H, W = p[1].shape
        want = [(B, PF_C, 2 * H, 2 * W), (B, PF_C, H, W), (B, PF_C, H // 2, W // 2), (B, PF_C, H // 4, W // 4)]
        if H % 4 or W % 4 or [tuple(t.shape) for t in p] != want:
            raise NotImplementedError('pf_semantic_fpn needs the four levels of a map whose sides are multiples of 4 (got %s)'
                                      % [tuple(t.shape) for t in p])
        HW = H * W
        HWp = round_up(HW, 8)
        fused = torch.empty((B, PF_C, HWp), dtype=torch.bfloat16, device=self.device)
        fused32 = torch.empty((B, PF_C, H, W), dtype=torch.float32, device=self.device) if want_fp32 else None
        nbytes = lib.pf_semantic_fpn_workspace_bytes(B, H, W)
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        _cabi.call('pf_semantic_fpn', ctypes.byref(self.struct), _ptr(p[0]), _ptr(p[1]), _ptr(p[2]), _ptr(p[3]), _ptr(fused),
                   _ptr(fused32), _ptr(self._ws), nbytes, B, H, W, HWp, _stream_ptr())
        self.last_launches = lib.pf_last_launch_count()
        return fused, fused32, (H, W)


class KernelHeadTail:
    """``KernelHead._decode_init_proposals`` from ``localization_feats`` on (kernel_head.py:250-347, eval mode,
    ``cat_stuff_mask=True``, ``use_binary=True``, ``proposal_feats_with_obj=True`` as shipped in
    configs/_base_/models/polyphonic_former.py:30-55)."""

    def __init__(self, state_dict, device, num_thing_classes=8):
        _cabi.load()
        self.device = torch.device(device)
        self.w = PackedKernelHead(state_dict, self.device, num_thing_classes)
        self._ws = None

    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws

    def cast_maps(self, localization_feats):
        """list of three fp32 [B,256,H,W] maps -> bf16 [3][B][256][HWp] (the storage dtype of the feature maps)."""
        B, C, H, W = localization_feats[0].shape
        HW = H * W
        HWp = round_up(HW, 8)
        out = torch.empty((3, B, PF_C, HWp), dtype=torch.bfloat16, device=self.device)
        for m, t in enumerate(localization_feats):
            t = t.to(self.device, torch.float32).contiguous()
            _cabi.call('pf_cast_maps', _ptr(t), _ptr(out[m]), B * PF_C, HW, HWp, _stream_ptr())
        return out

    def forward(self, maps, H, W, want_fp32_feats=False):
        """maps: bf16 [3][B][256][HWp] (cast_maps).  Returns a dict with the reference's 9-tuple members
        (kernel_head.py:347): proposal_feats [B,N,256,1,1], x_feats / depth_feats as ``feats`` (bf16 [2][B][256][HWp],
        the decoder's layout; fp32 copies under 'x_feats' / 'depth_feats' when asked), mask_preds [B,N,H,W],
        seg_preds [B,19,H,W], depth_proposal [B,N,256,1,1], depth_pred [B,1,H,W]; cls_scores and semantic_aspp_out
        are None in the reference's configuration."""
        lib = _cabi.load()
        w = self.w
        B, HWp = maps.shape[1], maps.shape[3]
        HW = H * W
        P, T = w.num_proposals, w.num_thing_classes
        N = P + w.num_classes - T
        dev, st = self.device, _stream_ptr()
        f32 = dict(dtype=torch.float32, device=dev)
        feats = torch.empty((2, B, PF_C, HWp), dtype=torch.bfloat16, device=dev)
        x32 = torch.empty((B, PF_C, H, W), **f32) if want_fp32_feats else None
        d32 = torch.empty((B, PF_C, H, W), **f32) if want_fp32_feats else None
        mask_preds = torch.empty((B, N, H, W), **f32)
        seg_preds = torch.empty((B, w.num_classes, H, W), **f32)
        depth_pred = torch.empty((B, 1, H, W), **f32)
        bits = torch.empty((B, (HW + 31) // 32, 128), dtype=torch.int32, device=dev)
        nbytes = lib.pf_kernel_head_workspace_bytes(B, HW)
        ws = self._workspace(nbytes)
        _cabi.call('pf_kernel_head', ctypes.byref(w.struct), _ptr(maps), _ptr(feats), _ptr(x32), _ptr(d32), _ptr(mask_preds),
                   _ptr(seg_preds), _ptr(depth_pred), _ptr(bits), _ptr(ws), nbytes, B, HW, HWp, st)
        # kernel_head.py:313-336: pool x_feats under the binarised initial masks, add init_kernels, append stuff kernels
        S = lib.pf_pool_splits(B, 1, HW)
        partial = torch.empty((B, S, P, PF_C), **f32)
        cntp = torch.empty((B, S, P), **f32)
        prop = torch.empty((B, N, PF_C), **f32)
        _cabi.call('pf_mask_pool', _ptr(feats), _ptr(bits), _ptr(partial), _ptr(cntp), B, P, HW, HWp, 1, S, st)
        _cabi.call('pf_init_proposals', _ptr(partial), _ptr(cntp), _ptr(w.init_kernels), _ptr(w.stuff_kernels),
                   _ptr(prop), B, P, N - P, S, st)
        self.last_launches = lib.pf_last_launch_count()   # counted since pf_kernel_head reset the counter
        dprop = w.depth_kernel.reshape(1, 1, PF_C, 1, 1).expand(B, N, PF_C, 1, 1)
        return dict(proposal_feats=prop.reshape(B, N, PF_C, 1, 1), feats=feats, x_feats=x32, depth_feats=d32,
                    mask_preds=mask_preds, cls_scores=None, seg_preds=seg_preds, depth_proposal=dprop,
                    depth_pred=depth_pred, semantic_aspp_out=None, bits=bits)
