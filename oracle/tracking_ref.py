"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch fp32) of the video model's tracking path, the next row
of the scope table, written BEFORE its kernels so that they have a pinned oracle:

  mask -> box                    polyphonic/video/utils.py:36-82 (RoI boxes), polyphonic/funcs/utils.py:4-22 (tracker boxes)
  RoI feature extraction         mmdet SingleRoIExtractor (single_level_roi_extractor.py:35-109) over mmcv RoIAlign
                                 (output 7x7, sampling_ratio 2, aligned=True, average pooling)
  embedding head                 polyphonic/video/track_heads.py:92-102 (4 x [3x3 conv + GN32 + ReLU], FC + ReLU, FC)
  association                    polyphonic/video/qdtrack/trackers/quasi_dense_embed_tracker.py:46-207
  the per-frame glue             polyphonic/polyphonic_former_video.py:364-403, 408-419

Pinned against the reference's own modules by tests/test_oracle_golden.py (fixture from oracle/make_golden.py: a
synthetic 4-frame clip).  Only tests/, smoke() and bench.py's cpu legs may import this module.
"""
import math

import torch
import torch.nn.functional as F
from torchvision.ops import roi_align


# ------------------------------------------------------------------------------------------------ mask -> box
def roi_box_of_mask(mask, extend=2.0):
    """video/utils.py:36-58 via batch_mask2boxlist :61-82: centre of the mask pixels +- extend x their mean absolute
    deviation (at least 1 pixel) per axis; [x1, y1, x2, y2]; an empty mask gives zeros."""
    ys, xs = torch.nonzero(mask, as_tuple=True)
    if ys.numel() == 0:
        return torch.zeros(4)
    ys, xs = ys.float(), xs.float()
    cy, cx = ys.mean(), xs.mean()
    dy = max((ys - cy).abs().mean(), torch.tensor(1.0))
    dx = max((xs - cx).abs().mean(), torch.tensor(1.0))
    return torch.stack([cx - dx * extend, cy - dy * extend, cx + dx * extend, cy + dy * extend])


def tight_box_of_mask(mask):
    """funcs/utils.py:4-22 (tensor_mask2box): the tight bounding box [x1, y1, x2, y2]; an empty mask gives (-1,-1,10,10)."""
    ys, xs = torch.nonzero(mask, as_tuple=True)
    if ys.numel() == 0:
        return torch.tensor([-1.0, -1.0, 10.0, 10.0])
    return torch.stack([xs.min(), ys.min(), xs.max(), ys.max()]).float()


# ------------------------------------------------------------------------------------------------ RoI features
def roi_features(feats, boxes, strides=(4, 8, 16, 32), out_size=7, sampling_ratio=2, finest_scale=56):
    """SingleRoIExtractor.forward for one image: each box goes to the pyramid level floor(log2(sqrt(w h) / 56 + 1e-6))
    clamped to [0, 3] and is RoIAligned there (aligned=True).  boxes [K,4] -> [K,256,7,7]."""
    K = boxes.shape[0]
    out = feats[0].new_zeros((K, feats[0].shape[1], out_size, out_size))
    if K == 0:
        return out
    rois = torch.cat([boxes.new_zeros((K, 1)), boxes], dim=1).clamp(min=0.0)        # polyphonic_former_video.py:414-415
    scale = torch.sqrt((rois[:, 3] - rois[:, 1]) * (rois[:, 4] - rois[:, 2]))
    lvl = torch.floor(torch.log2(scale / finest_scale + 1e-6)).clamp(min=0, max=len(strides) - 1).long()
    for i, s in enumerate(strides):
        sel = (lvl == i).nonzero(as_tuple=False).squeeze(1)
        if sel.numel():
            out[sel] = roi_align(feats[i], rois[sel], out_size, spatial_scale=1.0 / s, sampling_ratio=sampling_ratio,
                                 aligned=True)
    return out


def embed_head(sd, x, num_convs=4, num_groups=32, eps=1e-5):
    """QuasiDenseMaskEmbedHeadGTMask.forward (track_heads.py:92-102)."""
    for i in range(num_convs):
        x = F.conv2d(x, sd[f'convs.{i}.conv.weight'], padding=1)
        x = F.relu(F.group_norm(x, num_groups, sd[f'convs.{i}.gn.weight'], sd[f'convs.{i}.gn.bias'], eps))
    x = x.flatten(1)
    x = F.relu(F.linear(x, sd['fcs.0.weight'], sd['fcs.0.bias']))
    return F.linear(x, sd['fc_embed.weight'], sd['fc_embed.bias'])


def track_forward(sd, feats, masks):
    """PolyphonicVideo._track_forward (polyphonic_former_video.py:408-419) at test time: masks [K,H,W] of one frame."""
    boxes = torch.stack([roi_box_of_mask(m) for m in masks]) if len(masks) else torch.zeros((0, 4))
    return embed_head(sd, roi_features(feats, boxes))


# ------------------------------------------------------------------------------------------------ association
def pairwise_iou(a, b, eps=1e-6):
    """mmdet bbox_overlaps(mode='iou') for [M,4] x [N,4]."""
    if a.shape[0] * b.shape[0] == 0:
        return a.new_zeros((a.shape[0], b.shape[0]))
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    wh = (torch.min(a[:, None, 2:], b[None, :, 2:]) - torch.max(a[:, None, :2], b[None, :, :2])).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / torch.max(area_a[:, None] + area_b[None, :] - inter, inter.new_tensor([eps]))


class QuasiDenseTracker:
    """quasi_dense_embed_tracker.py with match_metric='bisoftmax', with_cats=True.  State: `tracks` (id -> box, embedding,
    label, last frame, mean velocity, frames accumulated) in insertion order, and the recent backdrops."""

    def __init__(self, init_score_thr=0.35, obj_score_thr=0.3, match_score_thr=0.5, memo_tracklet_frames=5,
                 memo_backdrop_frames=1, memo_momentum=0.8, nms_conf_thr=0.5, nms_backdrop_iou_thr=0.3,
                 nms_class_iou_thr=0.7):
        self.__dict__.update(locals())
        self.next_id = 0
        self.tracks = {}
        self.backdrops = []

    def _memory(self):
        """:103-135 -- tracklets first (insertion order), then the backdrops with id -1."""
        boxes = [t['box'][None] for t in self.tracks.values()] + [b['boxes'] for b in self.backdrops]
        embeds = [t['embed'][None] for t in self.tracks.values()] + [b['embeds'] for b in self.backdrops]
        labels = [t['label'].view(1) for t in self.tracks.values()] + [b['labels'] for b in self.backdrops]
        ids = list(self.tracks.keys()) + [-1] * sum(b['embeds'].shape[0] for b in self.backdrops)
        return torch.cat(boxes), torch.cat(labels), torch.cat(embeds), torch.tensor(ids, dtype=torch.long)

    def match(self, boxes, labels, embeds, frame_id):
        """:137-207.  boxes [K,5] (x1,y1,x2,y2,score).  Returns (boxes, labels, ids) of the detections that survive the
        duplicate removal, in descending score order."""
        order = boxes[:, 4].sort(descending=True)[1]
        boxes, labels, embeds = boxes[order], labels[order], embeds[order]
        iou = pairwise_iou(boxes[:, :4], boxes[:, :4])
        keep = torch.ones(boxes.shape[0], dtype=torch.bool)
        for i in range(1, boxes.shape[0]):          # every earlier box counts, suppressed or not (:147-153)
            thr = self.nms_backdrop_iou_thr if boxes[i, 4] < self.obj_score_thr else self.nms_class_iou_thr
            keep[i] = not bool((iou[i, :i] > thr).any())
        boxes, labels, embeds = boxes[keep], labels[keep], embeds[keep]
        ids = torch.full((boxes.shape[0],), -1, dtype=torch.long)
        if boxes.shape[0] > 0 and self.tracks:
            m_boxes, m_labels, m_embeds, m_ids = self._memory()
            sim = embeds @ m_embeds.t()
            score = (sim.softmax(dim=1) + sim.softmax(dim=0)) / 2
            score = score * (labels.view(-1, 1) == m_labels.view(1, -1)).float()
            for i in range(boxes.shape[0]):
                conf, j = score[i].max(dim=0)
                if conf > self.match_score_thr and m_ids[j] > -1:
                    if boxes[i, 4] > self.obj_score_thr:
                        ids[i] = m_ids[j]
                        score[:i, j] = 0
                        score[i + 1:, j] = 0
                    elif conf > self.nms_conf_thr:
                        ids[i] = -2
        fresh = (ids == -1) & (boxes[:, 4] > self.init_score_thr)
        n_new = int(fresh.sum())
        ids[fresh] = torch.arange(self.next_id, self.next_id + n_new, dtype=torch.long)
        self.next_id += n_new
        self._update(ids, boxes, embeds, labels, frame_id)
        return boxes, labels, ids

    def _update(self, ids, boxes, embeds, labels, frame_id):
        """:46-101."""
        for k in (ids > -1).nonzero(as_tuple=False).squeeze(1).tolist():
            tid, box, emb, lab = int(ids[k]), boxes[k], embeds[k], labels[k]
            t = self.tracks.get(tid)
            if t is None:
                self.tracks[tid] = dict(box=box, embed=emb, label=lab, last=frame_id, vel=torch.zeros_like(box), acc=0)
                continue
            vel = (box - t['box']) / (frame_id - t['last'])
            t['vel'] = (t['vel'] * t['acc'] + vel) / (t['acc'] + 1)
            t['acc'] += 1
            t['box'], t['label'], t['last'] = box, lab, frame_id
            t['embed'] = (1 - self.memo_momentum) * t['embed'] + self.memo_momentum * emb
        back = (ids == -1).nonzero(as_tuple=False).squeeze(1)
        iou = pairwise_iou(boxes[back, :4], boxes[:, :4])
        back = torch.tensor([int(b) for r, b in enumerate(back.tolist())
                             if not bool((iou[r, :b] > self.nms_backdrop_iou_thr).any())], dtype=torch.long)
        self.backdrops.insert(0, dict(boxes=boxes[back], embeds=embeds[back], labels=labels[back]))
        for tid in [k for k, t in self.tracks.items() if frame_id - t['last'] >= self.memo_tracklet_frames]:
            del self.tracks[tid]
        del self.backdrops[self.memo_backdrop_frames:]


def track_frame(sd, tracker, feats, masks, labels, scores, frame_id):
    """polyphonic_former_video.py:364-400 for one frame: thing masks [K,H,W] (bool) with their class labels and scores ->
    track ids (1-based; 0 = no track), aligned with the detections the tracker kept (descending score)."""
    if len(masks) == 0:
        return torch.zeros((0,), dtype=torch.long), None
    fm = masks.float()
    embeds = track_forward(sd, feats, fm)
    boxes = torch.cat([torch.stack([tight_box_of_mask(m) for m in fm]), scores.view(-1, 1).float()], dim=1)
    kept_boxes, kept_labels, ids = tracker.match(boxes, labels.long(), embeds, frame_id)
    ids = ids + 1
    ids[ids == -1] = 0
    return ids, kept_boxes
