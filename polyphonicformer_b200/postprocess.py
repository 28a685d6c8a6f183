"""Per-image post-processing of KernelUpdateIterHead.simple_test on the GPU:
``pf_panoptic`` replaces get_panoptic / merge_stuff_thing_stuff_joint (polyphonic/kernel_update.py:421-535) and
rescale_masks / rescale_depth (polyphonic/kernel_update_head.py:593-626).  Host code here only allocates, calls the C
ABI and turns the segment records into the reference's ``segments_info`` dicts.  No PyTorch fallback: geometries the
kernel does not cover raise."""
import ctypes

import numpy as np
import torch

from . import _cabi

_SEG_DTYPE = np.dtype([('id', np.int32), ('isthing', np.int32), ('category_id', np.int32), ('instance_id', np.int32),
                       ('area', np.int32), ('score', np.float32)])
_DEPTH_MODES = {'monodepth': 0, 'sigmoid': 1}


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def get_panoptic(roi_head, last_head, cls_scores, mask_preds, test_cfg, img_meta, depth_preds, depth_init,
                 aspp_semantic=None):
    """Same arguments and return value as the reference's get_panoptic: (None, None, (panoptic int32 [H0,W0],
    segments_info), depth_basic [H0,W0], depth_final [H0,W0]) as numpy."""
    if aspp_semantic is not None:
        raise NotImplementedError('aspp_semantic is not used by the reference either (kernel_update.py:425-426)')
    if not roi_head.merge_joint:
        raise NotImplementedError('merge_joint=False is not implemented by the reference (kernel_update.py:467)')
    if last_head.depth_act_mode not in _DEPTH_MODES:
        raise NotImplementedError('depth_act_mode=%r' % (last_head.depth_act_mode,))
    N, h, w = mask_preds.shape
    H0, W0 = img_meta['img_shape'][:2]
    Hb, Wb = img_meta['batch_input_shape']
    if (Hb, Wb) != (4 * h, 4 * w) or tuple(img_meta['ori_shape'][:2]) != (H0, W0):
        raise NotImplementedError('pf_panoptic covers predictions at 1/4 of the padded input and ori_shape == img_shape '
                                  '(got preds %dx%d, batch_input %dx%d, img %dx%d, ori %s); there is no PyTorch fallback'
                                  % (h, w, Hb, Wb, H0, W0, tuple(img_meta['ori_shape'][:2])))
    dev = mask_preds.device
    lib = _cabi.load()
    merge = test_cfg.merge_stuff_thing
    cls_scores = cls_scores.float().contiguous()
    mask_preds, depth_preds = mask_preds.float().contiguous(), depth_preds.float().contiguous()
    depth_init = depth_init.float().reshape(h, w).contiguous()
    pan = torch.empty((H0, W0), dtype=torch.int32, device=dev)
    dfinal = torch.empty((H0, W0), dtype=torch.float32, device=dev)
    dbasic = torch.empty((H0, W0), dtype=torch.float32, device=dev)
    segs = torch.zeros((128, _SEG_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    nseg = torch.zeros(1, dtype=torch.int32, device=dev)
    nbytes = lib.pf_panoptic_workspace_bytes(H0, W0)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _cabi.call('pf_panoptic', _ptr(cls_scores), _ptr(mask_preds), _ptr(depth_preds), _ptr(depth_init), N,
               roi_head.num_proposals, roi_head.num_thing_classes, cls_scores.shape[1], h, w, H0, W0,
               int(test_cfg.max_per_img), float(merge.instance_score_thr), float(merge.overlap_thr),
               _DEPTH_MODES[last_head.depth_act_mode], _ptr(pan), _ptr(dfinal), _ptr(dbasic), _ptr(segs), _ptr(nseg),
               _ptr(ws), nbytes, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    n = int(nseg.item())                                   # the one host synchronisation of the post-processing
    rec = np.frombuffer(segs.cpu().numpy().tobytes(), dtype=_SEG_DTYPE)[:n]
    info = []
    for r in rec:
        if r['isthing']:
            info.append({'id': int(r['id']), 'isthing': True, 'score': float(r['score']),
                         'category_id': int(r['category_id']), 'instance_id': int(r['instance_id'])})
        else:
            info.append({'id': int(r['id']), 'isthing': False, 'category_id': int(r['category_id']),
                         'area': int(r['area'])})
    return None, None, (pan.cpu().numpy(), info), dbasic.cpu().numpy(), dfinal.cpu().numpy()
