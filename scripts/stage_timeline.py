"""Debug (GPU): in-kernel timeline of the fused small-N stage kernel (csrc/pf_stage.cu): the 8 CTAs of unit 0, last launch.
Usage: python scripts/stage_timeline.py [B] [H] [W] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from polyphonicformer_b200 import _cabi  # noqa: E402

sys.argv = [sys.argv[0]] + (sys.argv[1:] or ['4', '128', '256', '3'])
tbuf = torch.zeros(16 + 16 * 16384, dtype=torch.int64, device='cuda:0')
_cabi.call('pf_debug_timeline', tbuf.data_ptr())
_cabi.load().pf_set_fused_update(1)
exec(open(os.path.join(ROOT, 'scripts', 'run_stage.py')).read())
torch.cuda.synchronize()
_cabi.call('pf_debug_timeline', None)
n = int(tbuf[0].item())
rec = tbuf[16:16 + 16 * n].reshape(n, 16).cpu()
steps = ['dual', 'gate', 'fc', 'qkv+att', 'out', 'ffn1a', 'ffn1b', 'ffn2a', 'ffn2b', 'reduce', 'heads', 'kern']
rows = {}
i = 0
while i + 5 < n:
    tag = int(rec[i, 15])
    if 500 <= tag < 508 and all(int(rec[i + j, 15]) == tag + 100 * j for j in range(6)):
        flat = [int(x) for j in range(6) for x in rec[i + j, :16]]
        rows.setdefault(tag - 500, []).append(flat)
        i += 6
    else:
        i += 1
def slot(v, k):
    return v[k + k // 15]
print('fused stage kernel, last launch, ns; per step: barrier->accumulators | epilogue part 1 | mid cluster barrier | part 2 | end barrier')
for r in sorted(rows):
    v = rows[r][-1]
    t0 = slot(v, 0)
    line = ['rank %d: pdl %d prep %d |' % (r, slot(v, 1) - t0, slot(v, 2) - slot(v, 1))]
    prev = slot(v, 2)
    for st, nm in enumerate(steps):
        b = 3 + 5 * st
        ts = [slot(v, b + j) for j in range(5)]
        seg = []
        p = prev
        for x in ts:
            seg.append(x - p if x else 0)
            p = x if x else p
        prev = p
        line.append('%s %s' % (nm, '/'.join(str(x) for x in seg)))
    line.append('| total %d' % (slot(v, 63) - t0))
    print('  '.join(line))
v = rows[0][-1]
pr = v[80:95]
base = slot(v, 3 + 5 * 1 + 4)   # end of the gates step
print('\nrank 0, fc step probes (ns after the barrier that ends the gates step):')
for nm, x in zip(['mma: full0', 'mma: full1', 'mma: full2', 'mma: full3', 'mma: committed', 'producer: start', 'producer: before empty wait (it=3)',
                  'producer: all loads issued'], pr[:8]):
    print('  %-36s %6d' % (nm, x - base))
