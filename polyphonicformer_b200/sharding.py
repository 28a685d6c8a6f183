"""Multi-GPU host logic.  One process per GPU; frames are independent units.

* Image mode: rank r decodes frames ``indices[r::world]`` (the reference's ``DistributedSampler``,
  ``mmdet/datasets/samplers/distributed_sampler.py:37``); weights are replicated; NO data-path collective.
* Video mode: the only cross-frame state of the reference is the tracker memo, and
  ``QuasiDenseEmbedTracker.match(bboxes[K,5], labels[K], track_feats[K,256], frame_id)``
  (``polyphonic/video/qdtrack/trackers/quasi_dense_embed_tracker.py:137-207``) consumes only those per-frame
  records.  Frames of a clip are dealt round-robin to the ranks, every rank runs backbone + decoder + track head on
  its frames, and ONE ``all_gather`` per clip of a fixed-size padded record tensor lets every rank replay the
  association in frame order -- identical to the reference's sequential loop
  (``polyphonic/apis/video_inference.py:8-37``).

Works with any ``torch.distributed`` backend: NCCL over NVLink for CUDA tensors (issue it on a side stream so it
overlaps the next clip's decoder), gloo for the CPU tests.
"""
import torch
import torch.distributed as dist

RECORD_WIDTH = 5 + 1 + 256   # bbox (x1, y1, x2, y2, score) | label | embedding


def shard_indices(n_items, rank, world):
    """Indices of the frames rank ``rank`` owns: ``range(n)[rank::world]``."""
    if not (0 <= rank < world):
        raise ValueError('rank %d outside world of %d' % (rank, world))
    return list(range(n_items))[rank::world]


def frames_per_rank(n_items, world):
    """Slots every rank must provide so that one fixed-size all_gather covers the clip."""
    return (n_items + world - 1) // world


def pack_records(records, slots, max_k, device=None):
    """records: list of (frame_id, bboxes [K,5], labels [K], feats [K,256]) -> float32 [slots, 2 + max_k*RECORD_WIDTH].
    Row layout: frame_id, K, then K records; unused slots have frame_id = -1.  K > max_k raises.  The headers of all
    slots go to the device in ONE copy; the bodies are device-to-device."""
    if len(records) > slots:
        raise ValueError('%d records for %d slots' % (len(records), slots))
    out = torch.zeros((slots, 2 + max_k * RECORD_WIDTH), dtype=torch.float32, device=device)
    head = torch.zeros((slots, 2), dtype=torch.float32)
    head[:, 0] = -1
    for i, (fid, bboxes, labels, feats) in enumerate(records):
        k = int(bboxes.shape[0])
        if k > max_k:
            raise ValueError('frame %d has %d tracks > max_k=%d' % (fid, k, max_k))
        head[i, 0], head[i, 1] = float(fid), float(k)
        if k:
            body = torch.cat([bboxes.reshape(k, 5).float(), labels.reshape(k, 1).float(), feats.reshape(k, 256).float()], 1)
            out[i, 2:2 + k * RECORD_WIDTH] = body.reshape(-1).to(out.device)
    out[:, :2] = head.to(out.device)
    return out


def unpack_records(packed, max_k):
    """Inverse of pack_records over the gathered [world*slots, ...] tensor; returns records sorted by frame id."""
    recs = []
    for row in packed:
        fid = int(row[0].item())
        if fid < 0:
            continue
        k = int(row[1].item())
        body = row[2:2 + k * RECORD_WIDTH].reshape(k, RECORD_WIDTH)
        recs.append((fid, body[:, :5].clone(), body[:, 5].long(), body[:, 6:].clone()))
    recs.sort(key=lambda r: r[0])
    return recs


def gather_frame_records(records, n_frames, max_k=100, group=None, device=None):
    """The ONE collective of the video path: every rank contributes the records of its frames, every rank gets all
    records of the clip in frame order.  ~105 KB per frame at max_k=100 (latency-bound, not bandwidth-bound)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    slots = frames_per_rank(n_frames, world)
    mine = pack_records(records, slots, max_k, device)
    if world == 1:
        return unpack_records(mine, max_k)
    gathered = torch.empty((world * slots, mine.shape[1]), dtype=mine.dtype, device=mine.device)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    return unpack_records(gathered.cpu(), max_k)
