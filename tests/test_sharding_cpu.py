"""N>1 host logic on CPU: world_size-2 gloo processes."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from polyphonicformer_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _clip(n_frames):
    """Deterministic per-frame records with ragged K (including K = 0)."""
    recs = []
    for f in range(n_frames):
        g = torch.Generator().manual_seed(100 + f)
        k = [3, 0, 7, 1, 5][f % 5]
        recs.append((f, torch.rand(k, 5, generator=g), torch.randint(0, 8, (k,), generator=g),
                     torch.randn(k, 256, generator=g)))
    return recs


def _worker(rank, world, port, n_frames, q):
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        clip = _clip(n_frames)
        mine = [clip[i] for i in sharding.shard_indices(n_frames, rank, world)]
        got = sharding.gather_frame_records(mine, n_frames, max_k=8)
        ok = len(got) == n_frames
        for (f, b, l, e), (f2, b2, l2, e2) in zip(got, clip):
            ok = ok and f == f2 and torch.equal(b, b2) and torch.equal(l, l2) and torch.equal(e, e2)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_shard_indices_partition_the_frames():
    for n in (0, 1, 5, 8, 37):
        for world in (1, 2, 4, 8):
            parts = [sharding.shard_indices(n, r, world) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) <= sharding.frames_per_rank(n, world)


def test_pack_unpack_round_trip_and_limits():
    clip = _clip(5)
    packed = sharding.pack_records(clip, 6, 8)
    assert packed.shape == (6, 2 + 8 * sharding.RECORD_WIDTH) and packed[5, 0] == -1
    back = sharding.unpack_records(packed, 8)
    assert [r[0] for r in back] == [0, 1, 2, 3, 4] and torch.equal(back[2][3], clip[2][3])
    import pytest
    with pytest.raises(ValueError):
        sharding.pack_records(clip, 6, 4)       # K = 7 > max_k
    with pytest.raises(ValueError):
        sharding.pack_records(clip, 4, 8)       # more frames than slots


def test_gather_frame_records_world2_gloo():
    """5-frame clip (config E shape) over 2 ranks: every rank ends with all 5 records in frame order."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 5, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]


def _tracking_worker(rank, world, port, q):
    """Video mode end to end on CPU: the frames of a clip are dealt to the ranks, every rank computes the per-frame
    records of ITS frames (here with the oracle standing in for the device path), one all_gather, then every rank
    replays the association in frame order."""
    import numpy as np
    from conftest import GOLDEN
    from oracle import synth, tracking_ref
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        gold = np.load(os.path.join(GOLDEN, 'tracking_clip_s0.npz'))
        clip = synth.synth_clip(seed=0)
        sd = synth.synth_track_head_state(0)
        mine = []
        with torch.no_grad():
            for f in sharding.shard_indices(len(clip), rank, world):
                fr = clip[f]
                fm = fr['masks'].float()
                boxes = torch.cat([torch.stack([tracking_ref.tight_box_of_mask(m) for m in fm]),
                                   fr['scores'].view(-1, 1).float()], dim=1)
                mine.append((f, boxes, fr['labels'], tracking_ref.track_forward(sd, fr['feats'], fm)))
        recs = sharding.gather_frame_records(mine, len(clip), max_k=8)
        tracker = tracking_ref.QuasiDenseTracker()
        ok = len(recs) == len(clip)
        for f, boxes, labels, embeds in recs:
            _, _, ids = tracker.match(boxes, labels, embeds, f + 1)
            ids = ids + 1
            ids[ids == -1] = 0
            ok = ok and ids.tolist() == gold[f'f{f}.ids'].tolist()
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_sharded_clip_tracking_matches_reference_world2_gloo():
    """Frame-sharded video mode reproduces the track ids of the reference's sequential loop (the golden of
    polyphonic_former_video.py:364-403 run on the same synthetic clip) on both ranks."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tracking_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == [(0, True), (1, True)]
