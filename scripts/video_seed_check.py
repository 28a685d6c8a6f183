"""Debug driver: bench.py's video step (1 rank) with the synthetic inputs rank `r` would use.  Usage: ... [r0] [r1]"""
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

r0 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
r1 = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device('cuda:0')
from polyphonicformer_b200.decoder import DecoderEngine  # noqa: E402
sd, stage_dicts = bench.synth_state()
eng = DecoderEngine(stage_dicts, dev, bench.NUM_CLASSES, 2048)
args = SimpleNamespace(steps=3)
for r in range(r0, r1):
    out = bench.run_video(args, eng, dev, 1, r, lambda: torch.cuda.synchronize())
    print('rank-seed', r, 'ok: things/frame', out['tracked_things_per_frame'], 'fps', round(out['value']), flush=True)
