"""Debug (GPU): timeline of the LAST decoder step: every pool / einsum CTA span and the fused stage kernel's unit-0 CTAs.
Usage: python scripts/step_timeline2.py [fused|per-layer] [B] [H] [W] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from polyphonicformer_b200 import _cabi  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'fused'
sys.argv = [sys.argv[0]] + (sys.argv[2:] or ['4', '128', '256', '3'])
tbuf = torch.zeros(16 + 16 * 32768, dtype=torch.int64, device='cuda:0')
_cabi.call('pf_debug_timeline', tbuf.data_ptr())
_cabi.load().pf_set_fused_update(1 if mode == 'fused' else 0)
exec(open(os.path.join(ROOT, 'scripts', 'run_stage.py')).read())
torch.cuda.synchronize()
_cabi.call('pf_debug_timeline', None)
n = int(tbuf[0].item())
rec = tbuf[16:16 + 16 * n].reshape(n, 16).cpu()
ev = []   # (start, end, name)
i = 0
while i < n:
    tag = int(rec[i, 15])
    if 500 <= tag < 508 and i + 5 < n and int(rec[i + 4, 15]) == tag + 400:
        flat = [int(x) for j in range(6) for x in rec[i + j, :16]]
        ev.append((flat[0], flat[63 + 63 // 15], 'stage(fused) r%d' % (tag - 500)))
        i += 6
        continue
    names = {1: 'prep', 2: 'sumln', 3: 'attention', 10: 'pool', 20: 'einsum(bits)', 21: 'einsum(logits)'}
    if tag in names or tag >= 1000:
        ev.append((int(rec[i, 0]), int(rec[i, 13]), names.get(tag, 'tcgemm %d' % tag)))
    i += 1
ev.sort()
# group consecutive events of the same name starting within 10 us into launches
launches = []
for st, en, nm in ev:
    base = nm.split(' r')[0]
    if launches and launches[-1]['name'] == base and st - launches[-1]['first'] < 10000:
        L = launches[-1]
        L['first_end'] = min(L['first_end'], en); L['end'] = max(L['end'], en); L['last_start'] = max(L['last_start'], st); L['n'] += 1
    else:
        launches.append(dict(name=base, first=st, last_start=st, first_end=en, end=en, n=1))
# the last step = everything after the last binarise-free gap: take the last `k` launches covering 3 pools
idx = [i for i, L in enumerate(launches) if L['name'] == 'pool'][-3]
last = launches[idx:]
t0 = last[0]['first']
print('%s path, last decoder step; ns relative to the first pool CTA' % mode)
print('%-22s %5s %9s %9s %9s %9s %8s' % ('kernel', 'ctas', 'start', 'laststart', 'firstend', 'end', 'gap'))
prev = None
for L in last:
    gap = L['first'] - prev if prev is not None else 0
    print('%-22s %5d %9d %9d %9d %9d %8d' % (L['name'], L['n'], L['first'] - t0, L['last_start'] - t0, L['first_end'] - t0, L['end'] - t0, gap))
    prev = L['end']
