"""N train-step-shaped calls of the drop-in modules (train.py:64-71) on BxHxW for profilers / sanitizers:
    python tools/modules_once.py 8x3072x4096 [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mmif_b200  # noqa: F401
from mmif_b200 import _lib as L
from mmif_b200.core import loss as ML
B, H, W = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else '8x3072x4096').split('x'))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
a, b, f = (torch.rand(B, 1, H, W, device='cuda') for _ in range(3))
fn1, fn2, fn3 = ML.SSIMLoss('ssim', weight=1.0), ML.PixelLoss('l1', weight=0.01), ML.GradLoss('l1', weight=0.1)
c0 = L.launch_counts()
for _ in range(steps):
    y = f.detach().requires_grad_(True)
    (fn1(a, b, y) + fn2(a, b, y, mode='max') + fn3(a, b, y, mode='max')).backward()
torch.cuda.synchronize()
c1 = L.launch_counts()
print('launches:', {k: c1[k] - c0[k] for k in c1}, 'loss vector', ML.last_loss_vector().tolist())
