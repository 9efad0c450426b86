"""Where the host time of one train-step-shaped loss call goes (8x1x256x256): cProfile over 3000 iterations."""
import cProfile, pstats, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mmif_b200  # noqa: F401
from mmif_b200.core import loss as ML
a, b, f = (torch.rand(8, 1, 256, 256, device='cuda') for _ in range(3))
fn1, fn2, fn3 = ML.SSIMLoss('ssim', weight=1.0), ML.PixelLoss('l1', weight=0.01), ML.GradLoss('l1', weight=0.1)


def step():
    y = f.detach().requires_grad_(True)
    (fn1(a, b, y) + fn2(a, b, y, mode='max') + fn3(a, b, y, mode='max')).backward()


for _ in range(200):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3000):
    step()
torch.cuda.synchronize()
print(f'wall per step: {(time.perf_counter() - t0) / 3000 * 1e6:.1f} us')
pr = cProfile.Profile()
pr.enable()
for _ in range(3000):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('tottime').print_stats(28)
