"""Who is right at full size?  Gradient of the fused objective on B x 3072 x 4096 uniform-random images: the drop-in
modules (fp32 kernels) and torch's fp32 CUDA eager graph (TF32 convolutions off) against torch's fp64 CUDA eager graph."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import mmif_b200  # noqa: F401
from mmif_b200.core import loss as ML

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(5)
x1, x2, y = (torch.rand(B, 1, 3072, 4096, device=dev, generator=g) for _ in range(3))


def grad_of(fn, a, b, f):
    f = f.detach().clone().requires_grad_(True)
    out = fn(a, b, f)
    sum(out).backward()
    return [o.item() for o in out], f.grad


fn1, fn2, fn3 = ML.SSIMLoss('ssim', weight=1.0), ML.PixelLoss('l1', weight=0.01), ML.GradLoss('l1', weight=0.1)
l_o, g_o = grad_of(lambda a, b, f: (fn1(a, b, f), fn2(a, b, f, mode='max'), fn3(a, b, f, mode='max')), x1, x2, y)
l_e, g_e = grad_of(bench.eager_objective, x1, x2, y)
l_d, g_d = grad_of(bench.eager_objective, x1.double(), x2.double(), y.double())
print('losses ours ', l_o)
print('losses eager', l_e)
print('losses fp64 ', l_d)
ref = g_d.abs().max().item()
for nm, gg in (('ours', g_o), ('eager fp32 (tf32 off)', g_e)):
    d = (gg.double() - g_d).abs()
    idx = torch.argmax(d).item()
    n, r, c = idx // (3072 * 4096), (idx // 4096) % 3072, idx % 4096
    frac = (d > 1e-5 * ref).double().mean().item()
    print(f'{nm}: max|g-g64|/max|g64| = {d.max().item() / ref:.3e} at (n={n}, row={r}, col={c}); fraction of elements off by > 1e-5 max|g64|: {frac:.3e}')
