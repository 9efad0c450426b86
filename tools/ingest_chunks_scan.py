"""Pipelined host ingest (core.metric.eval_metrics_batch_host): pairs/s against the number of chunks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mmif_b200  # noqa: F401
from mmif_b200.core import metric as MM

dev = torch.device('cuda:0')
for name, (n, h, w) in (('polar_32x1224x1024', (32, 1024, 1224)), ('tno_21x640x480', (21, 480, 640))):
    g = torch.Generator().manual_seed(7)
    a = torch.randint(0, 256, (n, 1, h, w), generator=g, dtype=torch.uint8)
    b = torch.randint(0, 256, (n, 1, h, w), generator=g, dtype=torch.uint8)
    f = ((a.int() + b.int()) // 2).to(torch.uint8)
    hu = [t.pin_memory() for t in (a, b, f)]
    hf = [t.float().pin_memory() for t in (a, b, f)]
    rows_host = torch.empty(n, 16, dtype=torch.float64).pin_memory()
    for label, src in (('u8', hu), ('f32', hf)):
        line = []
        for chunks in (1, 2, 3, 4, 6, 8, 16):
            def run():
                rows_host.copy_(MM.eval_metrics_batch_host(*src, chunks=chunks), non_blocking=True)
                torch.cuda.synchronize()
            for _ in range(3):
                run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(8):
                run()
            e1.record(); torch.cuda.synchronize()
            line.append(f'{chunks}:{n / (e0.elapsed_time(e1) / 8 * 1e-3):.0f}')
        print(name, label, 'pairs/s by chunks', ' '.join(line))
