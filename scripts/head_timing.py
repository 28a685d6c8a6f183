"""Device time of the KernelHead tail (pf_kernel_head + pf_mask_pool + pf_init_proposals) at BASELINE.json's
1024x2048 configuration (decoder map 128x256), per launch, with CUDA events on the launching stream."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth
from polyphonicformer_b200.kernel_head import KernelHeadTail

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H, W = 128, 256
dev = torch.device('cuda:0')
tail = KernelHeadTail(synth.synth_kernel_head_state(0), dev)
sets = []
for i in range(3):   # rotating input sets: 3 x 201 MB of bf16 maps > L2
    maps = torch.relu(torch.randn(3, B, 256, H * W, device=dev)).to(torch.bfloat16)
    sets.append(maps)
for i in range(3):
    tail.forward(sets[i % 3], H, W)
torch.cuda.synchronize()
n = 20
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for i in range(n):
    tail.forward(sets[i % 3], H, W)
ev[1].record()
torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / n
HW = H * W
alg = B * HW * (3 * 256 * 2 + 2 * 256 * 2 + (111 + 19 + 1) * 4)   # maps in, feats out, predictions out (bytes)
print(json.dumps({'B': B, 'ms_per_call': ms, 'frames_per_s': B / ms * 1e3, 'launches': tail.last_launches,
                  'algorithmic_MB': alg / 1e6, 'algorithmic_GBps': alg / ms / 1e6}))
