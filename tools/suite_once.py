"""Run the batched metric suite a few times (for ncu launch lists): python tools/suite_once.py N H W iters"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mmif_b200
from mmif_b200.core import metric as MM
n, h, w, iters = (int(v) for v in (sys.argv[1:5] + ['32', '1024', '1224', '3'][len(sys.argv) - 1:]))
g = torch.Generator(device='cuda').manual_seed(7)
a = torch.randint(0, 256, (n, 1, h, w), device='cuda', generator=g).float()
b = torch.randint(0, 256, (n, 1, h, w), device='cuda', generator=g).float()
f = torch.floor((a + b) / 2)
for _ in range(iters):
    MM.eval_metrics_batch(a, b, f)
torch.cuda.synchronize()
print('done')
