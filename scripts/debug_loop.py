"""Debug helper (GPU): chain stage_forward without teacher forcing and report where the fused loop departs from the golden."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from conftest import GOLDEN, rel_err
from oracle import synth
from polyphonicformer_b200.decoder import DecoderEngine

dev = torch.device('cuda:0')
for name in ['decoder_b2_h16_w24_s6', 'decoder_b1_h10_w12_s1']:
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    B, H, W, seed = int(g['B']), int(g['H']), int(g['W']), int(g['seed'])
    sd = synth.synth_decoder_state(3, seed)
    stage_dicts = [{k[len('mask_head.%d.' % s):]: v for k, v in sd.items() if k.startswith('mask_head.%d.' % s)} for s in range(3)]
    eng = DecoderEngine(stage_dicts, dev)
    inp = synth.synth_decoder_inputs(B, H, W, seed)
    feats = eng.prepare_feats(inp['x_feats'].to(dev), inp['depth_feats'].to(dev))
    mask, obj, dep = inp['mask_preds'].to(dev), inp['proposal_feats'].reshape(B, -1, 256).to(dev), inp['depth_proposal'].reshape(B, -1, 256).to(dev)
    print('==', name)
    for s in range(3):
        cls, logits, obj, dep = eng.stage_forward(s, feats, mask, obj, dep, H, W, cls_sigmoid=(s == 2))
        torch.cuda.synchronize()
        mask = logits[0]
        gm = torch.from_numpy(g['s%d.mask_preds' % s])
        fl = ((mask.cpu() > 0) != (gm > 0))
        print(' stage', s, 'mask', rel_err(mask.cpu(), gm), 'depth', rel_err(logits[1].cpu(), g['s%d.depth_preds' % s]),
              'obj', rel_err(obj.cpu().flatten(), g['s%d.object_feats' % s].flatten()), 'flips', int(fl.sum()),
              'min|logit|', gm.abs().min().item())
        if fl.any():
            idx = fl.nonzero()
            print('   flipped at', idx[:5].tolist(), 'golden vals', gm[fl][:5].tolist(), 'mask area of those rows',
                  [(gm[i[0], i[1]] > 0).sum().item() for i in idx[:5]])
    gc = torch.from_numpy(g['cls_score_sigmoid'])
    d = (cls.cpu() - gc).abs()
    print(' cls', rel_err(cls.cpu(), gc), 'worst rows', d.amax(-1).flatten().topk(3))
