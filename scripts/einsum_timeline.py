import os, sys, torch
ROOT='/root/repo'; sys.path.insert(0, ROOT)
from polyphonicformer_b200 import _cabi
sys.argv=[sys.argv[0],'4','128','256','3']
tbuf=torch.zeros(16+16*16384,dtype=torch.int64,device='cuda:0')
_cabi.call('pf_debug_timeline', tbuf.data_ptr())
exec(open(os.path.join(ROOT,'scripts','run_stage.py')).read())
torch.cuda.synchronize(); _cabi.call('pf_debug_timeline', None)
n=int(tbuf[0].item()); rec=tbuf[16:16+16*n].reshape(n,16).cpu()
r=rec[(rec[:,15]==20)&(rec[:,14]==0)]
for x in r[-2:]:
    t0=int(x[0]); print('einsum<0> CTA0:', [int(v)-t0 if v else 0 for v in x[:14]])
r=rec[(rec[:,15]==21)&(rec[:,14]==0)]
for x in r[-1:]:
    t0=int(x[0]); print('einsum<1> CTA0:', [int(v)-t0 if v else 0 for v in x[:14]])
