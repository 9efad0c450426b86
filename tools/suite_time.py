"""Event-timed throughput of the batched metric suite on the two BASELINE shapes (not the contract bench)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mmif_b200
from mmif_b200.core import metric as MM
for (n, h, w) in ((21, 480, 640), (32, 1024, 1224), (1, 1024, 1224)):
    g = torch.Generator(device='cuda').manual_seed(7)
    a = torch.randint(0, 256, (n, 1, h, w), device='cuda', generator=g).float()
    b = torch.randint(0, 256, (n, 1, h, w), device='cuda', generator=g).float()
    f = torch.floor((a + b) / 2)
    for _ in range(3):
        MM.eval_metrics_batch(a, b, f)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        MM.eval_metrics_batch(a, b, f)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f'{n}x{h}x{w}: {ms:.3f} ms/batch  {n / ms * 1e3:.0f} pairs/s  {87.6 * n * h * w / ms / 1e6:.0f} GB/s @87.6 B/px')
