"""Host side of the video model's tracking path (include/pf_track.h): the per-frame glue of
``PolyphonicVideo.simple_test`` after the panoptic merge (polyphonic/polyphonic_former_video.py:364-451 of the reference)
on the CUDA kernels of pf_track.cu.  PyTorch is device memory and streams only; there is no fallback.

  TrackHeadEngine   mask -> box, RoIAlign and QuasiDenseMaskEmbedHeadGTMask   (:408-419, video/utils.py:40-82,
                    video/track_heads.py:92-102)
  DeviceTracker     QuasiDenseEmbedTracker with the memo on the device         (qdtrack/trackers/quasi_dense_embed_tracker.py)

State-dict keys consumed (prefix ``track_head.`` in a full model): convs.{0..3}.conv.weight [256,256,3,3],
convs.{0..3}.gn.{weight,bias} [256], fcs.0.{weight [1024,12544], bias}, fc_embed.{weight [256,1024], bias}.
"""
import ctypes

import numpy as np
import torch

from . import _cabi
from ._cabi import TrackerConfig, TrackWeights

MAX_K = 128
EMBED = 256


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _hi_lo(w32):
    hi = w32.to(torch.bfloat16)
    lo = (w32 - hi.to(torch.float32)).to(torch.bfloat16)
    return hi, lo


class PackedTrackHead:
    """QuasiDenseMaskEmbedHeadGTMask parameters in the layouts ``struct pf_track_weights`` documents."""

    def __init__(self, sd, device, gn_eps=1e-5):
        g = lambda k: sd[k].detach().to('cpu', torch.float32)
        n_convs = len([k for k in sd if k.startswith('convs.') and k.endswith('.conv.weight')])
        if n_convs != 4:
            raise NotImplementedError('pf_track_embed is built for num_convs=4 (got %d)' % n_convs)
        planes = []
        for l in range(4):
            w = g(f'convs.{l}.conv.weight')                               # [out, in, ky, kx]
            if tuple(w.shape) != (256, 256, 3, 3):
                raise NotImplementedError('track head conv %d has shape %s, expected 256x256x3x3' % (l, tuple(w.shape)))
            taps = w.permute(2, 3, 0, 1).reshape(9, 256, 256)             # [ky*3+kx][out][in]
            hi, lo = _hi_lo(taps)
            planes += [hi, lo]
        self.conv_w = torch.stack(planes).contiguous().to(device)         # [4*2][9][256][256]
        self.gn_gamma = torch.stack([g(f'convs.{l}.gn.weight') for l in range(4)]).contiguous().to(device)
        self.gn_beta = torch.stack([g(f'convs.{l}.gn.bias') for l in range(4)]).contiguous().to(device)
        fc1 = g('fcs.0.weight')
        if tuple(fc1.shape) != (1024, 256 * 49) or tuple(sd['fc_embed.weight'].shape) != (EMBED, 1024):
            raise NotImplementedError('track head FC shapes %s / %s' % (tuple(fc1.shape), tuple(sd['fc_embed.weight'].shape)))
        fc1 = fc1.reshape(1024, 256, 49).permute(0, 2, 1).reshape(1024, 49 * 256)     # input index (c, y, x) -> (y, x, c)
        hi, lo = _hi_lo(fc1)
        self.fc1_w = torch.stack([hi, lo]).contiguous().to(device)        # [2][1024][12544]
        self.fc1_b = g('fcs.0.bias').contiguous().to(device)
        self.fc2_wt = g('fc_embed.weight').t().contiguous().to(device)    # [1024][256]
        self.fc2_b = g('fc_embed.bias').contiguous().to(device)
        self.struct = TrackWeights(conv_w=self.conv_w.data_ptr(), gn_gamma=self.gn_gamma.data_ptr(),
                                   gn_beta=self.gn_beta.data_ptr(), fc1_w=self.fc1_w.data_ptr(), fc1_b=self.fc1_b.data_ptr(),
                                   fc2_wt=self.fc2_wt.data_ptr(), fc2_b=self.fc2_b.data_ptr(), gn_eps=gn_eps)


class TrackHeadEngine:
    """mask -> box, RoI features and embeddings of one frame's thing masks."""

    def __init__(self, state_dict, device, strides=(4, 8, 16, 32)):
        self.device = torch.device(device)
        self.lib = _cabi.load()
        if self.device.type != 'cuda':
            raise _cabi.PFError(-4, 'TrackHeadEngine', 'the tracking path needs a CUDA (sm_100) device; there is no CPU fallback')
        self.packed = PackedTrackHead(state_dict, self.device)
        self.strides = tuple(int(s) for s in strides)
        if len(self.strides) != 4:
            raise NotImplementedError('pf_track_embed takes the 4 FPN levels of the shipped configs')

    def _boxes(self, fn, src_args, K, H, W):
        rois = torch.empty((K, 5), dtype=torch.float32, device=self.device)
        tight = torch.empty((K, 4), dtype=torch.float32, device=self.device)
        nbytes = self.lib.pf_track_boxes_workspace_bytes(K, H, W)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        _cabi.call(fn, *src_args, K, H, W, _ptr(rois), _ptr(tight), _ptr(ws), nbytes, _stream())
        return rois, tight

    def boxes_from_masks(self, masks):
        """masks [K,H,W] (any dtype, non-zero = inside) -> (rois [K,5] for the RoI extractor, tight boxes [K,4])."""
        masks = masks.to(self.device, torch.float32).contiguous()
        K, H, W = masks.shape
        return self._boxes('pf_track_boxes_from_masks', (_ptr(masks),), K, H, W)

    def boxes_from_panoptic(self, panoptic, seg_ids):
        """panoptic int32 [H,W] on the device, seg_ids: the segment ids of the K thing segments."""
        H, W = panoptic.shape
        ids = torch.as_tensor(seg_ids, dtype=torch.int32, device=self.device)
        return self._boxes('pf_track_boxes_from_panoptic', (_ptr(panoptic), _ptr(ids)), int(ids.numel()), H, W)

    def embed(self, feats, rois, want_roi_feats=False):
        """feats: the FPN levels [1,256,h_l,w_l] fp32 of ONE image; rois [K,5] -> embeddings [K,256]."""
        K = int(rois.shape[0])
        lv = [f.reshape(f.shape[-3], f.shape[-2], f.shape[-1]).to(self.device, torch.float32).contiguous() for f in feats[:4]]
        if len(lv) != 4 or any(f.shape[0] != 256 for f in lv):
            raise NotImplementedError('pf_track_embed needs 4 levels of 256 channels')
        ptrs = (ctypes.c_void_p * 4)(*[f.data_ptr() for f in lv])
        hs = (ctypes.c_int * 4)(*[f.shape[1] for f in lv])
        wss = (ctypes.c_int * 4)(*[f.shape[2] for f in lv])
        st = (ctypes.c_int * 4)(*self.strides)
        emb = torch.empty((K, EMBED), dtype=torch.float32, device=self.device)
        rf = torch.empty((K, 256, 7, 7), dtype=torch.float32, device=self.device) if want_roi_feats else None
        nbytes = self.lib.pf_track_embed_workspace_bytes(K)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        _cabi.call('pf_track_embed', ctypes.byref(self.packed.struct), ptrs, hs, wss, st, _ptr(rois.contiguous()), K, _ptr(emb),
                   _ptr(rf), _ptr(ws), nbytes, _stream())
        return (emb, rf) if want_roi_feats else emb

    def head(self, roi_feats):
        """QuasiDenseMaskEmbedHeadGTMask.forward: RoI features [K,256,7,7] -> embeddings [K,256]."""
        x = roi_feats.to(self.device, torch.float32).contiguous()
        K = int(x.shape[0])
        emb = torch.empty((K, EMBED), dtype=torch.float32, device=self.device)
        nbytes = self.lib.pf_track_embed_workspace_bytes(K)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        _cabi.call('pf_track_head', ctypes.byref(self.packed.struct), _ptr(x), K, _ptr(emb), _ptr(ws), nbytes, _stream())
        return emb


class DeviceTracker:
    """QuasiDenseEmbedTracker (bisoftmax, with_cats) with its memo in device memory; ``match`` has the reference's
    signature and return value (kept boxes, labels, ids in descending score order)."""

    def __init__(self, device, init_score_thr=0.8, obj_score_thr=0.5, match_score_thr=0.5, memo_tracklet_frames=10,
                 memo_backdrop_frames=1, memo_momentum=0.8, nms_conf_thr=0.5, nms_backdrop_iou_thr=0.3,
                 nms_class_iou_thr=0.7, with_cats=True, match_metric='bisoftmax'):
        if match_metric != 'bisoftmax':
            raise NotImplementedError('match_metric=%r (the shipped configs use bisoftmax)' % match_metric)
        assert 0 <= memo_momentum <= 1.0 and memo_tracklet_frames >= 0 and memo_backdrop_frames >= 0
        self.device = torch.device(device)
        self.lib = _cabi.load()
        if self.device.type != 'cuda':
            raise _cabi.PFError(-4, 'DeviceTracker', 'the tracker needs a CUDA (sm_100) device; there is no CPU fallback')
        self.cfg = TrackerConfig(init_score_thr=init_score_thr, obj_score_thr=obj_score_thr, match_score_thr=match_score_thr,
                                 memo_momentum=memo_momentum, nms_conf_thr=nms_conf_thr,
                                 nms_backdrop_iou_thr=nms_backdrop_iou_thr, nms_class_iou_thr=nms_class_iou_thr,
                                 memo_tracklet_frames=memo_tracklet_frames, memo_backdrop_frames=memo_backdrop_frames,
                                 with_cats=int(bool(with_cats)))
        self.state = torch.empty(self.lib.pf_tracker_state_bytes(), dtype=torch.uint8, device=self.device)
        self.ws_bytes = self.lib.pf_tracker_workspace_bytes()
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        self.out = torch.zeros(2 * MAX_K + 2, dtype=torch.int32, device=self.device)      # order | ids | n_kept | status
        self.reset()

    def reset(self):
        _cabi.call('pf_tracker_reset', _ptr(self.state), _stream())

    def match_async(self, bboxes, labels, track_feats, frame_id):
        """Launch only.  Returns the int32 device tensor [order(128) | ids(128) | n_kept | status] of this frame (a view of a
        buffer the next call overwrites: clone it to keep it)."""
        K = int(bboxes.shape[0])
        b = bboxes.to(self.device, torch.float32).contiguous()
        l = labels.to(self.device, torch.int32).contiguous()
        e = track_feats.to(self.device, torch.float32).contiguous()
        if e.data_ptr() % 16:          # a one-row view of a larger buffer counts as contiguous whatever its offset
            e = e.clone()
        o = self.out
        _cabi.call('pf_tracker_match', ctypes.byref(self.cfg), _ptr(self.state), _ptr(b), _ptr(l), _ptr(e), K, int(frame_id),
                   _ptr(o), _ptr(o[MAX_K:]), _ptr(o[2 * MAX_K:]), _ptr(o[2 * MAX_K + 1:]), _ptr(self.ws), self.ws_bytes,
                   _stream())
        return o

    def match(self, bboxes, labels, track_feats, frame_id):
        """quasi_dense_embed_tracker.py:137-207.  Returns (bboxes [n,5], labels [n], ids [n] int64) on the inputs' device."""
        o = self.match_async(bboxes, labels, track_feats, frame_id).cpu()
        n = int(o[2 * MAX_K])
        if int(o[2 * MAX_K + 1]):
            raise _cabi.PFError(-1, 'pf_tracker_match', 'tracklet memo overflow (> %d live tracks)' % 512)
        order = o[:n].long()
        ids = o[MAX_K:MAX_K + n].long()
        order_d = order.to(bboxes.device)
        return bboxes[order_d], labels[order_d], ids


def paint_maps_batch(panoptic_devs, segments_infos, seg_ids_list, ids_list, default_sem, host_buffer=None):
    """generate_track_id_maps + get_semantic_seg (polyphonic_former_video.py:436-451) for several frames of one shape: two
    look-up tables per frame over its panoptic map (pf_track_paint), ONE upload of the tables and ONE read-back of the maps.
    As in the reference, ids[i] (kept detections, descending score) is painted onto the i-th thing mask in segment order;
    pixels of no segment get ``default_sem`` / track id 0.  Returns per frame numpy (sem uint8, track float64 -- the
    reference's np.zeros(shape))."""
    F = len(panoptic_devs)
    if F == 0:
        return []
    dev, shape = panoptic_devs[0].device, tuple(panoptic_devs[0].shape)
    n = panoptic_devs[0].numel()
    luts = np.zeros((F, 1280), dtype=np.uint8)                     # per frame: track lut (256 x int32) | sem lut (256 x uint8)
    for f in range(F):
        trk_lut = np.zeros(256, dtype=np.int32)
        for i, tid in enumerate(ids_list[f]):
            trk_lut[seg_ids_list[f][i]] = int(tid)
        luts[f, :1024] = trk_lut.view(np.uint8)
        luts[f, 1024:] = default_sem
        for s in segments_infos[f]:
            luts[f, 1024 + s['id']] = s['category_id']
    luts_d = torch.from_numpy(luts).to(dev)
    out = torch.empty(F * n * 9, dtype=torch.uint8, device=dev)      # all track maps (float64), then all semantic maps (uint8)
    trk = out[:F * n * 8].view(torch.float64).view(F, n)
    sem = out[F * n * 8:].view(F, n)
    for f in range(F):
        _cabi.call('pf_track_paint', _ptr(panoptic_devs[f]), _ptr(luts_d[f, 1024:]), _ptr(luts_d[f]), n, _ptr(sem[f]), _ptr(trk[f]),
                   1, _stream())
    if host_buffer is not None and host_buffer.numel() >= out.numel():
        host_t = host_buffer[:out.numel()]          # the caller's pinned buffer: the returned arrays are views of it
        host_t.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        host = host_t.numpy()
    else:
        host = out.cpu().numpy()
    trk_h = host[:F * n * 8].view(np.float64).reshape((F,) + shape)
    sem_h = host[F * n * 8:].reshape((F,) + shape)
    return [(sem_h[f], trk_h[f]) for f in range(F)]


def paint_maps(panoptic_dev, segments_info, seg_ids, ids, default_sem):
    """One frame of paint_maps_batch."""
    return paint_maps_batch([panoptic_dev], [segments_info], [seg_ids], [ids], default_sem)[0]
