"""Quick device-side timing of the fused loss kernels (not the contract benchmark; see bench.py)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mmif_b200
from mmif_b200 import _lib as L
from mmif_b200.core import loss as ML

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

shapes = [(8, 3072, 4096), (1, 1024, 1224), (8, 256, 256), (64, 256, 256)]
if len(sys.argv) > 1: shapes = [tuple(int(v) for v in sys.argv[1].split('x'))]
lib = L.load()
for (B, H, W) in shapes:
    a, b, f = (torch.rand(B, 1, H, W, device='cuda') for _ in range(3))
    cfg = ML._cfg(1.0, 'max', 'max', 'l1', 'l1')
    out = torch.empty(lib.mmif_loss_out_doubles(B), dtype=torch.float64, device='cuda')
    ws = torch.zeros(lib.mmif_loss_workspace_bytes(B, H, W), dtype=torch.uint8, device='cuda')
    g = torch.ones(3, device='cuda'); dF = torch.empty_like(f)
    st = L.stream_ptr(a.device)
    fwd = lambda: L.check(lib.mmif_fusion_loss_fwd(a.data_ptr(), b.data_ptr(), f.data_ptr(), B, H, W, ctypes.byref(cfg), out.data_ptr(), None, ws.data_ptr(), ws.numel(), st))
    bwd = lambda: L.check(lib.mmif_fusion_loss_bwd(a.data_ptr(), b.data_ptr(), f.data_ptr(), B, H, W, ctypes.byref(cfg), g.data_ptr(), None, dF.data_ptr(), ws.data_ptr(), ws.numel(), st))
    cfgz = ML._cfg(1.0, 'max', 'max', 'l1', 'l1'); cfgz.want_grad = 1
    dU = torch.empty_like(f)
    zfwd = lambda: L.check(lib.mmif_fusion_loss_fwd(a.data_ptr(), b.data_ptr(), f.data_ptr(), B, H, W, ctypes.byref(cfgz), out.data_ptr(), dU.data_ptr(), ws.data_ptr(), ws.numel(), st))
    zbwd = lambda: L.check(lib.mmif_fusion_loss_bwd(a.data_ptr(), b.data_ptr(), f.data_ptr(), B, H, W, ctypes.byref(cfg), g.data_ptr(), dU.data_ptr(), dF.data_ptr(), ws.data_ptr(), ws.numel(), st))
    tf, tb = timeit(fwd), timeit(bwd)
    tzf, tzb = timeit(zfwd), timeit(zbwd)
    mp = B * H * W / 1e6
    print(f'{B}x{H}x{W}: fwd {tf:.3f} ms ({mp/tf/1e3*1e3:.1f} Mpix/ms = {mp/tf:.0f} Gpix/s*1e-3) bwd {tb:.3f} ms  fwd+bwd {mp/(tf+tb)*1e3:.0f} Mpix/s  '
          f'GB/s bwd {16*mp/tb:.0f} both {28*mp/(tf+tb):.0f} | single-pass: fwd+grad {tzf:.3f} ms rescale {tzb:.3f} ms -> {mp/(tzf+tzb)*1e3:.0f} Mpix/s ({16*mp/tzf:.0f} GB/s @16B/px)')
