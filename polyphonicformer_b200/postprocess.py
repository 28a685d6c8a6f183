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


class _PinnedPool:
    """Read-back buffers for results that are handed to the caller as numpy arrays.  A fresh pageable array per call costs
    more than the whole decoder (100 MB of page faults + a staged copy at a fraction of the link rate), so results are read
    into PINNED buffers that the returned arrays own: ``lend`` gives out a tensor view and its numpy array, and when the last
    array derived from that array is garbage-collected the buffer goes back to the pool.  A caller that keeps every result (mmdet's test loop does) just makes
    the pool grow up to ``max_outstanding`` buffers; past that, results fall back to ordinary pageable memory."""

    def __init__(self, max_outstanding=8):
        self.free, self.outstanding, self.max_outstanding = {}, 0, max_outstanding

    def lend(self, nbytes):
        import weakref
        size = 1 << max(20, (nbytes - 1).bit_length())          # power-of-two classes
        bucket = self.free.setdefault(size, [])
        if bucket:
            base = bucket.pop()
        elif self.outstanding < self.max_outstanding:
            base = torch.empty(size, dtype=torch.uint8).pin_memory()
        else:
            return None
        self.outstanding += 1
        view = base[:nbytes]
        host = view.numpy()                                     # every slice / .view() of it keeps `host` alive through .base
        weakref.finalize(host, self._release, size, base)
        return view, host

    def _release(self, size, base):
        self.outstanding -= 1
        self.free.setdefault(size, []).append(base)


_POOL = _PinnedPool()


def _check_geometry(h, w, img_meta):
    H0, W0 = img_meta['img_shape'][:2]
    Hb, Wb = img_meta['batch_input_shape']
    if (Hb, Wb) != (4 * h, 4 * w) or tuple(img_meta['ori_shape'][:2]) != (H0, W0):
        raise NotImplementedError('pf_panoptic covers predictions at 1/4 of the padded input and ori_shape == img_shape '
                                  '(got preds %dx%d, batch_input %dx%d, img %dx%d, ori %s); there is no PyTorch fallback'
                                  % (h, w, Hb, Wb, H0, W0, tuple(img_meta['ori_shape'][:2])))
    return H0, W0


def _segments_info(rec):
    info = []
    for r in rec:
        if r['isthing']:
            info.append({'id': int(r['id']), 'isthing': True, 'score': float(r['score']),
                         'category_id': int(r['category_id']), 'instance_id': int(r['instance_id'])})
        else:
            info.append({'id': int(r['id']), 'isthing': False, 'category_id': int(r['category_id']),
                         'area': int(r['area'])})
    return info


def get_panoptic_batch(roi_head, last_head, cls_scores, mask_preds, test_cfg, img_metas, depth_preds, depth_init,
                       stride2_inputs=False, device_results=None, host_buffer=None):
    """get_panoptic (kernel_update.py:421-469) for the B frames of a batch in ONE set of launches and ONE device->host
    copy.  cls_scores [B,N,classes]; mask_preds / depth_preds [B,N,h,w]; depth_init [B,h,w] (or [B,1,h,w]).

    ``stride2_inputs``: the three maps are the decoder's OWN stride-8 outputs [.., h/2, w/2] and the x2 bilinear
    up-sampling of kernel_update.py:131-143 / :302-307 is evaluated inside the kernels (same arithmetic, bit-identical
    results), so scaled_mask_preds / scaled_depth_preds never exist.  Frames must share one img_shape (a batch does:
    kernel_update.py:339-351 loops over frames of one padded batch); mixed shapes go frame by frame.
    Returns the reference's list of (None, None, (panoptic, segments_info), depth_basic, depth_final), numpy.
    ``device_results``: an optional list that receives, per frame, a dict of the DEVICE tensors behind those arrays
    (panoptic int32 [H0,W0], depth_final, depth_basic) -- the video model's tracking path keeps working on them.
    ``host_buffer``: an optional PINNED uint8 tensor the results are read back into (the returned arrays are then views of
    it and live as long as the caller keeps it untouched); by default a fresh pageable array per call."""
    if not roi_head.merge_joint:
        raise NotImplementedError('merge_joint=False is not implemented by the reference (kernel_update.py:467)')
    if last_head.depth_act_mode not in _DEPTH_MODES:
        raise NotImplementedError('depth_act_mode=%r' % (last_head.depth_act_mode,))
    B, N = mask_preds.shape[:2]
    k = 2 if stride2_inputs else 1
    h, w = k * mask_preds.shape[2], k * mask_preds.shape[3]
    shapes = {_check_geometry(h, w, m) for m in img_metas}
    if len(img_metas) != B:
        raise ValueError('%d img_metas for a batch of %d' % (len(img_metas), B))
    if len(shapes) > 1:      # ragged crops inside one padded batch: one frame at a time, same kernels
        out = []
        for b in range(B):
            out += get_panoptic_batch(roi_head, last_head, cls_scores[b:b + 1], mask_preds[b:b + 1], test_cfg,
                                      img_metas[b:b + 1], depth_preds[b:b + 1], depth_init[b:b + 1], stride2_inputs,
                                      device_results)
        return out
    H0, W0 = shapes.pop()
    dev = mask_preds.device
    lib = _cabi.load()
    merge = test_cfg.merge_stuff_thing
    cls_scores = cls_scores.float().contiguous()
    mask_preds, depth_preds = mask_preds.float().contiguous(), depth_preds.float().contiguous()
    depth_init = depth_init.float().reshape(B, h // k, w // k).contiguous()
    # one output allocation = one read-back: panoptic | depth_final | depth_basic | segment records | counts
    npx = H0 * W0
    seg_bytes = 128 * _SEG_DTYPE.itemsize
    offs = np.cumsum([0, B * npx * 4, B * npx * 4, B * npx * 4, B * seg_bytes, B * 4])
    out = torch.zeros(int(offs[-1]), dtype=torch.uint8, device=dev)
    pan, dfinal, dbasic, segs, nseg = (out[int(a):int(b)] for a, b in zip(offs[:-1], offs[1:]))
    nbytes = B * lib.pf_panoptic_workspace_bytes(H0, W0)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _cabi.call('pf_panoptic_batch', _ptr(cls_scores), _ptr(mask_preds), _ptr(depth_preds), _ptr(depth_init), B, N,
               roi_head.num_proposals, roi_head.num_thing_classes, cls_scores.shape[-1], h, w, H0, W0,
               int(test_cfg.max_per_img), float(merge.instance_score_thr), float(merge.overlap_thr),
               _DEPTH_MODES[last_head.depth_act_mode], 1 if stride2_inputs else 0, _ptr(pan), _ptr(dfinal), _ptr(dbasic),
               _ptr(segs), 128, _ptr(nseg), _ptr(ws), nbytes, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    if device_results is not None:
        pd, fd, bd = (t.view(torch.int32 if i == 0 else torch.float32).view(B, H0, W0) for i, t in enumerate((pan, dfinal, dbasic)))
        device_results += [dict(panoptic=pd[b], depth_final=fd[b], depth_basic=bd[b]) for b in range(B)]
    if host_buffer is not None and host_buffer.numel() >= out.numel():
        host_t = host_buffer[:out.numel()]
        host_t.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        host = host_t.numpy()
    else:
        lent = _POOL.lend(out.numel())
        if lent is not None:
            host_t, host = lent                             # the returned arrays keep `host` (and the pinned buffer) alive
            host_t.copy_(out, non_blocking=True)
            torch.cuda.current_stream().synchronize()       # the one host synchronisation of the post-processing
        else:
            host = out.cpu().numpy()
    pan_h = host[offs[0]:offs[1]].view(np.int32).reshape(B, H0, W0)
    dfinal_h = host[offs[1]:offs[2]].view(np.float32).reshape(B, H0, W0)
    dbasic_h = host[offs[2]:offs[3]].view(np.float32).reshape(B, H0, W0)
    rec = host[offs[3]:offs[4]].view(_SEG_DTYPE).reshape(B, 128)
    n = host[offs[4]:offs[5]].view(np.int32)
    return [(None, None, (pan_h[b], _segments_info(rec[b, :n[b]])), dbasic_h[b], dfinal_h[b]) for b in range(B)]


def get_panoptic(roi_head, last_head, cls_scores, mask_preds, test_cfg, img_meta, depth_preds, depth_init,
                 aspp_semantic=None):
    """Same arguments and return value as the reference's get_panoptic: (None, None, (panoptic int32 [H0,W0],
    segments_info), depth_basic [H0,W0], depth_final [H0,W0]) as numpy."""
    if aspp_semantic is not None:
        raise NotImplementedError('aspp_semantic is not used by the reference either (kernel_update.py:425-426)')
    N, h, w = mask_preds.shape
    return get_panoptic_batch(roi_head, last_head, cls_scores[None], mask_preds[None], test_cfg, [img_meta],
                              depth_preds[None], depth_init.reshape(1, h, w))[0]
