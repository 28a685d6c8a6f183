"""Debug (GPU): per-kernel device times of pf_panoptic on one 1024x2048 frame (ncu-free: CUDA events per call)."""
import ctypes, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth
from polyphonicformer_b200 import _cabi
dev = torch.device('cuda:0')
h, w = 256, 512
H0, W0 = 4 * h, 4 * w
inp = {k: v.to(dev) for k, v in synth.synth_panoptic_inputs(h, w, 0).items()}
lib = _cabi.load()
P = lambda t: ctypes.c_void_p(t.data_ptr())
pan = torch.empty((H0, W0), dtype=torch.int32, device=dev)
df, db = torch.empty((H0, W0), device=dev), torch.empty((H0, W0), device=dev)
segs = torch.zeros((128, 24), dtype=torch.uint8, device=dev)
nseg = torch.zeros(1, dtype=torch.int32, device=dev)
nb = lib.pf_panoptic_workspace_bytes(H0, W0)
ws = torch.empty(nb, dtype=torch.uint8, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def call():
    _cabi.call('pf_panoptic', P(inp['cls_scores']), P(inp['mask_preds']), P(inp['depth_preds']), P(inp['depth_init']), 111, 100, 8, 19,
               h, w, H0, W0, 100, 0.3, 0.6, 1, P(pan), P(df), P(db), P(segs), P(nseg), P(ws), nb, st)
for _ in range(3): call()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): call()
b.record(); torch.cuda.synchronize()
print('pf_panoptic device time per frame: %.3f ms' % (a.elapsed_time(b) / 20), 'segments', int(nseg.item()))
