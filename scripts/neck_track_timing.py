"""Profiling driver (GPU) for the kernels added in round 2, nothing else on the device:
  neck      pf_semantic_fpn + pf_fpn_pred on the four FPN levels of B frames (decoder map 128x256)
  track     pf_track_boxes_from_panoptic + pf_track_embed + pf_tracker_match on 30 RoIs of a 1024x2048 frame
  panoptic  pf_panoptic_batch on the decoder's stride-8 logits of B frames
Usage: python scripts/neck_track_timing.py neck|track|panoptic [B] [reps]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402  (weights / synthetic inputs only)

what = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device('cuda:0')
H, W = 128, 256
g = torch.Generator().manual_seed(0)
if what == 'neck':
    from polyphonicformer_b200.kernel_head import FpnPred, SemanticFpnPyramid
    sd = synth.synth_semantic_fpn_state(0)
    pyr, pred = SemanticFpnPyramid(sd, dev), FpnPred(sd, dev)
    levels = [torch.randn(B, 256, 2 * H >> i, 2 * W >> i, generator=g).to(dev) for i in range(4)]

    def run():
        fused, _, hw = pyr.forward(levels)
        pred.forward(fused, hw=hw)
elif what == 'track':
    from polyphonicformer_b200.track import DeviceTracker, TrackHeadEngine
    trk = TrackHeadEngine(synth.synth_track_head_state(0), dev)
    tracker = DeviceTracker(dev, init_score_thr=0.35, obj_score_thr=0.3, memo_tracklet_frames=5)
    K = 30
    pan = torch.zeros((8 * H, 8 * W), dtype=torch.int32)
    for k in range(K):
        pan[(k // 6) * 200 + 5:(k // 6) * 200 + 150, (k % 6) * 340 + 10:(k % 6) * 340 + 300] = k + 1
    pan = pan.to(dev)
    fpn = [torch.randn(256, 2 * H >> i, 2 * W >> i, generator=g).to(dev) for i in range(4)]
    labels = torch.arange(K, device=dev) % 8
    scores = torch.linspace(0.95, 0.4, K, device=dev).view(-1, 1)
    frame = [0]

    def run():
        frame[0] += 1
        rois, tight = trk.boxes_from_panoptic(pan, list(range(1, K + 1)))
        emb = trk.embed(fpn, rois)
        tracker.match_async(torch.cat([tight, scores], 1), labels, emb, frame[0])
else:
    import ctypes
    from types import SimpleNamespace
    from polyphonicformer_b200 import postprocess
    from polyphonicformer_b200.registry import to_config
    cfg = to_config(json.load(open(os.path.join(ROOT, 'tests', 'golden', 'roi_head_cfg.json')))['test_cfg'])
    roi = SimpleNamespace(num_proposals=100, num_thing_classes=8, merge_joint=True)
    last = SimpleNamespace(depth_act_mode='sigmoid', num_classes=19)
    fr = [synth.synth_panoptic_inputs(H, W, s) for s in range(B)]
    cls = torch.stack([f['cls_scores'] for f in fr]).to(dev)
    mask = torch.stack([f['mask_preds'] for f in fr]).to(dev)
    depth = torch.stack([f['depth_preds'] for f in fr]).to(dev)
    dinit = torch.stack([f['depth_init'] for f in fr]).to(dev)
    meta = dict(img_shape=(8 * H, 8 * W, 3), ori_shape=(8 * H, 8 * W, 3), batch_input_shape=(8 * H, 8 * W))

    def run():
        postprocess.get_panoptic_batch(roi, last, cls, mask, cfg, [meta] * B, depth, dinit, stride2_inputs=True)

for _ in range(2):
    run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    run()
b.record()
torch.cuda.synchronize()
print(json.dumps({'what': what, 'B': B, 'ms_per_call': a.elapsed_time(b) / reps}))
