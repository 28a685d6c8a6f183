"""Debug (GPU): one line per kernel launch of the last decoder step: first CTA start, last CTA end, gap to previous."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from polyphonicformer_b200 import _cabi  # noqa: E402

sys.argv = [sys.argv[0]] + (sys.argv[1:] or ['4', '128', '256', '3'])
tbuf = torch.zeros(16 + 16 * 16384, dtype=torch.int64, device='cuda:0')
_cabi.call('pf_debug_timeline', tbuf.data_ptr())
exec(open(os.path.join(ROOT, 'scripts', 'run_stage.py')).read())
torch.cuda.synchronize()
_cabi.call('pf_debug_timeline', None)
n = int(tbuf[0].item())
rec = tbuf[16:16 + 16 * n].reshape(n, 16).cpu()
rec = rec[rec[:, 0].argsort()]
# group records into launches: same tag, starts within 8 us
launches = []
for r in rec:
    tag, st, en = int(r[15]), int(r[0]), int(r[13])
    if launches and launches[-1]['tag'] == tag and st - launches[-1]['first'] < 8000 and tag in (10, 20, 21):
        L = launches[-1]
        L['end'] = max(L['end'], en); L['n'] += 1
    else:
        launches.append(dict(tag=tag, first=st, end=en, n=1))
per_step = len(launches) // steps
last = launches[-per_step:]
names = {1: 'prep', 2: 'sumln', 3: 'attention', 10: 'pool', 20: 'einsum(bits)', 21: 'einsum(logits)'}
t0 = last[0]['first']
prev_end = None
print('%-22s %10s %10s %8s %8s' % ('kernel', 'start', 'end', 'dur', 'gap'))
for L in last:
    nm = names.get(L['tag'], 'tcgemm %d' % L['tag'])
    gap = (L['first'] - prev_end) if prev_end is not None else 0
    print('%-22s %10d %10d %8d %8d' % (nm, L['first'] - t0, L['end'] - t0, L['end'] - L['first'], gap))
    prev_end = L['end']
