"""Debug (GPU): in-kernel timeline of the small-N block's GEMM launches (CTA 0 of each), in ns relative to the first."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from polyphonicformer_b200 import _cabi  # noqa: E402

sys.argv = [sys.argv[0]] + (sys.argv[1:] or ['4', '128', '256', '3'])
tbuf = torch.zeros(16 + 16 * 4096, dtype=torch.int64, device='cuda:0')
_cabi.call('pf_debug_timeline', tbuf.data_ptr())
exec(open(os.path.join(ROOT, 'scripts', 'run_stage.py')).read())
torch.cuda.synchronize()
_cabi.call('pf_debug_timeline', None)
n = int(tbuf[0].item())
rec = tbuf[16:16 + 16 * n].reshape(n, 16).cpu()
rec = rec[(rec[:, 15] < 10) | (rec[:, 15] >= 100)]   # small-N block only (pool / einsum record every CTA)
rec = rec[rec[:, 0].argsort()]
last = rec[-36:]           # the last step: 12 launches per stage
t0 = int(last[0, 0])
names = ['start', 'setup', 'pdlwait', 'full0', 'full1', 'full2', 'full3', '-', 'accfull', 'phase1', 'stats', 'cbar', 'done', 'exit']
print('tag    ' + ' '.join(n.rjust(8) for n in names))
for r in last:
    print(str(int(r[15])).ljust(6), ' '.join((str(int(v) - t0) if v else '-').rjust(8) for v in r[:14]))
