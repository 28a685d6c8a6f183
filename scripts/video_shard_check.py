"""Frame-sharded video mode on hardware: run under torchrun with N ranks (NCCL).  Every rank takes
its round-robin share of two 4-frame synthetic clips (+ a frame without things), runs the tracking head on the device,
the ONE all_gather of the per-frame records, replays the association and paints its frames; rank 0 also runs the whole
sequence through oracle/tracking_ref.py and every rank's frames are compared with it (track ids exact, maps identical).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/video_shard_check.py
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    from oracle import synth, tracking_ref
    from polyphonicformer_b200.track import TrackHeadEngine
    from polyphonicformer_b200.video import VideoShardRunner
    import test_video_gpu as tv
    cfg = tv.video_cfg()
    sd = synth.synth_track_head_state(0)
    frames = tv.clip_frames(seed=0) + tv.clip_frames(seed=1) + tv.clip_frames(seed=2) + tv.clip_frames(seed=3)
    n_local = len(frames) // world
    frames = frames[:n_local * world]
    runner = VideoShardRunner(SimpleNamespace(device=dev), TrackHeadEngine(sd, dev), None, None, None, cfg['tracker'], tv.NUM_THING,
                              tv.NUM_STUFF, clip_len=4)
    tcfg = {k: v for k, v in cfg['tracker'].items() if k not in ('type', 'with_cats', 'match_metric')}
    want, ref, cnt = [], None, 1
    for g, fr in enumerate(frames):
        if g % 4 == 0:
            ref, cnt = tracking_ref.QuasiDenseTracker(**tcfg), 1
        want.append(tv.expected_maps(ref, sd, fr, cnt))
        cnt += 1
    gids = runner.global_ids(0, n_local)
    local_frames = [frames[g] for g in gids]
    res = [(None, None, (fr['panoptic'].numpy(), fr['info']), None, fr['depth']) for fr in local_frames]
    dev_res = [dict(panoptic=fr['panoptic'].to(dev)) for fr in local_frames]
    fpn = [torch.stack([fr['feats'][l][0] for fr in local_frames]).to(dev) for l in range(4)]
    recs = runner.track_records(res, dev_res, fpn)
    ids = runner.associate(recs, gids)
    out = runner.paint(dev_res, res, recs, gids, ids)
    ok = True
    for g, o in zip(gids, out):
        ok &= ids.get(g, []) == want[g][0] and np.array_equal(o['track'], want[g][1]) and np.array_equal(o['sem'], want[g][2])
    flag = torch.tensor([int(ok)], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print('rank %d/%d frames %s ids %s -> %s' % (rank, world, gids, [ids.get(g, []) for g in gids], 'OK' if ok else 'MISMATCH'), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print('VIDEO_SHARD_CHECK %s world=%d backend=%s' % ('OK' if int(flag) else 'FAILED', world, 'nccl' if world > 1 else 'none'))
    sys.exit(0 if int(flag) else 1)


if __name__ == '__main__':
    main()
