"""Frame-sharded video inference (video mode; BASELINE.json configs 3 and 5).

The reference processes a clip strictly frame by frame on one GPU (polyphonic/apis/video_inference.py:8-37): decoder,
panoptic merge, tracking head, then ``QuasiDenseEmbedTracker.match`` against the memo of the previous frames
(polyphonic/polyphonic_former_video.py:326-403).  Only that last step depends on the other frames, and only through one
small record per frame -- ``(bboxes [K,5], labels [K], track_feats [K,256])``.  So the frames of a wave are dealt round
robin to the ranks (global frame g -> rank g % world), every rank runs everything up to the records for ITS frames as one
batch, ONE ``all_gather`` (NCCL over NVLink; gloo in the CPU tests) brings all records to every rank, and every rank
replays the association in frame order on the device -- ``pf_tracker_match`` per frame, the memo reset at clip boundaries
-- which reproduces the reference's sequential loop exactly.  Each rank then paints the track-id / semantic maps of its
own frames.
"""
import torch
import torch.distributed as dist

from . import postprocess, sharding
import time

from .track import MAX_K, DeviceTracker, paint_maps_batch


class VideoShardRunner:
    """One rank's share of the video path from the decoder's inputs on.

    decoder: DecoderEngine; track: TrackHeadEngine; ``roi_head`` / ``last_head`` / ``test_cfg``: what
    postprocess.get_panoptic_batch needs (num_proposals, num_thing_classes, merge_joint / depth_act_mode / the rcnn test
    config); tracker_cfg: the kwargs of QuasiDenseEmbedTracker."""

    def __init__(self, decoder, track, roi_head, last_head, test_cfg, tracker_cfg, num_thing_classes, num_stuff_classes,
                 clip_len, group=None, max_k=100):
        self.dec, self.trk = decoder, track
        self.roi_head, self.last_head, self.test_cfg = roi_head, last_head, test_cfg
        self.num_thing_classes, self.num_stuff_classes = num_thing_classes, num_stuff_classes
        self.clip_len, self.group, self.max_k = int(clip_len), group, int(max_k)
        self.device = decoder.device
        self.tracker = DeviceTracker(self.device, **{k: v for k, v in tracker_cfg.items() if k != 'type'})
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.cnt = 1
        self.comm_stream = torch.cuda.Stream(self.device)
        self._bufs = {}           # decoder output buffers and pinned read-back buffers, per batch shape
        self.timing = None        # set to a dict to collect per-section host times (seconds, synchronised) in step()

    def global_ids(self, wave, n_local):
        """Global frame numbers of this rank's n_local frames of wave `wave` (a wave = world * n_local consecutive frames)."""
        base = wave * self.world * n_local
        return [base + i * self.world + self.rank for i in range(n_local)]

    def decode_panoptic(self, batch, H, W):
        """Decoder + batched panoptic merge for this rank's frames.  Returns (the reference's per-frame result tuples, the
        device tensors behind them)."""
        F, N = batch['mask'].shape[:2]
        key = (F, N, H, W)
        if key not in self._bufs:
            npx = 64 * H * W
            self._bufs[key] = dict(dec=self.dec.alloc_decode_buffers(F, N, H, W, upsample=False),
                                   pan=torch.empty(F * (npx * 12 + 128 * 24 + 4), dtype=torch.uint8).pin_memory(),
                                   paint=torch.empty(F * npx * 9, dtype=torch.uint8).pin_memory())
        bufs = self._bufs[key]
        out = self.dec.decode(batch['feats'], batch['mask'], batch['prop'], batch['dprop'], H, W, upsample=False,
                              buffers=bufs['dec'])
        dev_res = []
        res = postprocess.get_panoptic_batch(self.roi_head, self.last_head, out['cls_score'], out['mask_preds'], self.test_cfg,
                                             batch['img_metas'], out['depth_preds'], batch['depth_pred'],
                                             stride2_inputs=True, device_results=dev_res, host_buffer=bufs['pan'])
        return res, dev_res

    def track_records(self, res, dev_res, fpn):
        """polyphonic_former_video.py:364-390 without the memo: per frame (seg_ids, bboxes [K,5], labels [K], embeds [K,256])
        on the device, or None for a frame without thing segments.  fpn: the 4 levels [F,256,h_l,w_l]."""
        recs = []
        for f in range(len(res)):
            things = [s for s in res[f][2][1] if s['isthing']]
            if not things:
                recs.append(None)
                continue
            seg_ids = [s['id'] for s in things]
            rois, tight = self.trk.boxes_from_panoptic(dev_res[f]['panoptic'], seg_ids)
            embeds = self.trk.embed([lv[f] for lv in fpn], rois)
            scores = torch.tensor([s['score'] for s in things], dtype=torch.float32, device=self.device)
            labels = torch.tensor([s['category_id'] for s in things], dtype=torch.int64, device=self.device)
            recs.append((seg_ids, torch.cat([tight, scores.view(-1, 1)], 1), labels, embeds))
        return recs

    def _mark(self, name, t0):
        if self.timing is not None:
            torch.cuda.synchronize()
            self.timing[name] = self.timing.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    def step(self, batch, H, W, wave):
        """One wave: this rank's frames end to end.  Returns the reference's per-frame dicts (sem, track, depth); sem and
        depth are views of per-runner pinned buffers the next step overwrites (so are track maps)."""
        F, N = batch['mask'].shape[:2]
        gids = self.global_ids(wave, F)
        t = time.perf_counter()
        res, dev_res = self.decode_panoptic(batch, H, W)
        t = self._mark('decode_panoptic', t)
        recs = self.track_records(res, dev_res, batch['fpn'])
        t = self._mark('track_records', t)
        ids = self.associate(recs, gids)
        t = self._mark('associate', t)
        out = self.paint(dev_res, res, recs, gids, ids, host_buffer=self._bufs[(F, N, H, W)]['paint'])
        self._mark('paint', t)
        return out

    def associate(self, recs, gids):
        """The ONE collective + the replay.  Returns {global frame id: track ids (+1, 0 = none) of its kept detections}."""
        mine = [(g, r[1], r[2], r[3]) if r is not None else
                (g, torch.zeros((0, 5), device=self.device), torch.zeros((0,), device=self.device),
                 torch.zeros((0, 256), device=self.device)) for g, r in zip(gids, recs)]
        n_wave = self.world * len(gids)
        base = min(gids) - self.rank if gids else 0
        slots = len(gids)
        packed = sharding.pack_records([(g - base, b, l, e) for g, b, l, e in mine], slots, self.max_k, self.device)
        if self.world > 1:
            # on a side stream: the collective is ordered after the records, not after whatever else the main stream holds
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                gathered = torch.empty((self.world * slots, packed.shape[1]), dtype=packed.dtype, device=self.device)
                dist.all_gather_into_tensor(gathered, packed, group=self.group)
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        else:
            gathered = packed
        # frame order = sorted by frame id; the header (frame id, K) of every row comes to the host in one copy
        head = gathered[:, :2].cpu()
        order = sorted((int(head[i, 0]), i) for i in range(head.shape[0]) if head[i, 0] >= 0)
        assert len(order) == n_wave, (len(order), n_wave)
        outs = {}
        for rel, row in order:
            g = base + rel
            if g % self.clip_len == 0:                       # video_inference.py:24-25: a new clip re-creates the tracker
                self.tracker.reset()
                self.cnt = 1
            k = int(head[row, 1])
            if k == 0:
                continue                                     # polyphonic_former_video.py:371, 400: no match(), cnt stays
            body = gathered[row, 2:2 + k * sharding.RECORD_WIDTH].view(k, sharding.RECORD_WIDTH)
            o = self.tracker.match_async(body[:, :5], body[:, 5], body[:, 6:], self.cnt)
            self.cnt += 1
            if g in gids:
                outs[g] = o.clone()
        result = {}
        if outs:
            keys = sorted(outs)
            host = torch.stack([outs[g] for g in keys]).cpu()      # the one synchronisation of the replay
            for g, o in zip(keys, host):
                if int(o[2 * MAX_K + 1]):
                    raise RuntimeError('tracklet memo overflow in frame %d' % g)
                ids = o[MAX_K:MAX_K + int(o[2 * MAX_K])].long() + 1
                ids[ids == -1] = 0                                  # polyphonic_former_video.py:398-399
                result[g] = ids.tolist()
        return result

    def paint(self, dev_res, res, recs, gids, ids_by_frame, host_buffer=None):
        maps = paint_maps_batch([d['panoptic'] for d in dev_res], [r[2][1] for r in res],
                                [rec[0] if rec is not None else [] for rec in recs], [ids_by_frame.get(g, []) for g in gids],
                                self.num_thing_classes + self.num_stuff_classes, host_buffer)
        return [{'sem': sem, 'track': trk, 'depth': r[4]} for (sem, trk), r in zip(maps, res)]
