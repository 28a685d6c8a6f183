"""Profiling driver (GPU): a few decoder steps at a given shape, nothing else.
Usage: python scripts/run_stage.py [B] [H] [W] [steps]   (H, W = decoder map size, i.e. frame size / 8)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402  (weights only; nothing from oracle/ is executed on the measured path)
from polyphonicformer_b200.decoder import DecoderEngine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H = int(sys.argv[2]) if len(sys.argv) > 2 else 128
W = int(sys.argv[3]) if len(sys.argv) > 3 else 256
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = torch.device('cuda:0')
sd = synth.synth_decoder_state(3, 0)
stage_dicts = [{k[len('mask_head.%d.' % s):]: v for k, v in sd.items() if k.startswith('mask_head.%d.' % s)}
               for s in range(3)]
eng = DecoderEngine(stage_dicts, dev)
g = torch.Generator().manual_seed(0)
x = torch.relu(torch.randn(B, 256, H, W, generator=g)).to(torch.bfloat16).to(dev)
d = torch.relu(torch.randn(B, 256, H, W, generator=g)).to(torch.bfloat16).to(dev)
mask = torch.randn(B, 111, H, W, generator=g).to(dev)
prop = (torch.randn(B, 111, 256, generator=g) * 0.5).to(dev)
dprop = (torch.randn(B, 111, 256, generator=g) * 0.1).to(dev)
feats = eng.prepare_feats(x, d)
buf = eng.alloc_decode_buffers(B, 111, H, W, upsample=True)
for _ in range(steps):
    buf['obj'].copy_(prop), buf['dep'].copy_(dprop)
    eng.decode_inplace(feats, mask, buf, H, W)
torch.cuda.synchronize()
print('ok')
