"""Debug (GPU): per-CTA start / end times of the pooling and einsum launches of the last decoder step."""
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
for tag in (10, 20, 21):
    r = rec[rec[:, 15] == tag]
    if not len(r):
        continue
    # group into launches by time gaps: records of one launch start within a few us of each other
    order = r[:, 0].argsort()
    r = r[order]
    launches, cur = [], [r[0]]
    for x in r[1:]:
        if int(x[0]) - int(cur[-1][0]) > 8000:
            launches.append(torch.stack(cur)); cur = []
        cur.append(x)
    launches.append(torch.stack(cur))
    L = launches[-1]
    t0 = int(L[:, 0].min())
    st, en = L[:, 0] - t0, L[:, 13] - t0
    print('tag %d: %d CTAs; start min/median/max %d/%d/%d ns; end min/median/max %d/%d/%d ns; tile4 median %d; last issue median %d'
          % (tag, len(L), st.min(), st.median(), st.max(), en.min(), en.median(), en.max(), (L[:, 2] - t0).median(),
             (L[:, 3] - t0).median()))
    dur = (L[:, 13] - L[:, 0]).float()
    print('   per-CTA duration min/median/max %.1f/%.1f/%.1f us' % (dur.min() / 1e3, dur.median() / 1e3, dur.max() / 1e3))
