"""One small pass through every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize_once.py

Shapes are small (the tools slow kernels down 10-100x) but chosen so that every path runs: interior and
border batches / strips of the loss kernels (single-pass, recomputing, forward-only), a width TMA cannot
describe (plain-load ring), the metric suite with its pyramids and the packed-counter histogram.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import mmif_b200  # noqa: F401,E402
from mmif_b200.core import loss as ML  # noqa: E402
from mmif_b200.core import metric as MM  # noqa: E402

g = torch.Generator().manual_seed(3)
for (B, H, W) in ((2, 96, 352), (1, 70, 131)):          # 352: 4 strips, interior ones take the fast Sobel path; 131: no TMA
    a, b, f = (torch.rand(B, 1, H, W, generator=g).cuda() for _ in range(3))
    f.requires_grad_(True)
    l1, l2, l3 = ML.SSIMLoss('ssim')(a, b, f), ML.PixelLoss('l1', 0.01)(a, b, f, mode='max'), ML.GradLoss('l1', 0.1)(a, b, f, mode='max')
    (l1 + l2 + l3).backward(retain_graph=True)           # single-pass + in-place backward
    torch.autograd.grad(2.0 * l1 + l3, f)                 # recomputing backward
    with torch.no_grad():
        ML.SSIMLoss('ssim')(a, b, f.detach()).item()      # forward-only kernel
    f3 = f.detach().clone().requires_grad_(True)          # avg / l2 modes: the general instantiation of the warp-specialised kernel
    (ML.SSIMLoss('ssim')(a, b, f3) + ML.PixelLoss('l2', 0.01)(a, b, f3, mode='avg') + ML.GradLoss('l2', 0.1)(a, b, f3, mode='avg')).backward()
    f2 = f.detach().clone().requires_grad_(True)
    ML.SSIMLoss('ms-ssim')(torch.rand(1, 1, 192, 208, generator=g).cuda(), torch.rand(1, 1, 192, 208, generator=g).cuda(),
                           torch.rand(1, 1, 192, 208, generator=g).cuda().requires_grad_(True)).backward()
    ML.SSIMLoss('msw-ssim')(a, b, f2).backward()
for (n, h, w) in ((2, 200, 264), (1, 97, 131)):
    x, y = (torch.randint(0, 256, (n, 1, h, w), generator=g).float().cuda() for _ in range(2))
    z = torch.floor((x + y) / 2)
    MM.eval_metrics_batch(x, y, z)
torch.cuda.synchronize()
print('sanitize_once: done')
