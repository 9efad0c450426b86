"""Scan forced row-segment geometries of the warp-specialised kernel (MMIF_WS_GEOM): python tools/ws_geom_force.py B H W"""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import mmif_b200
from mmif_b200 import _lib as L
from mmif_b200.core import loss as ML
B, H, W = (int(v) for v in sys.argv[1:4])
lib = L.load()
a, b, f = (torch.rand(B, 1, H, W, device='cuda') for _ in range(3))
st = L.stream_ptr(a.device)
cfgz = ML._cfg(1.0, 'max', 'max', 'l1', 'l1', 1.0, 0.01, 0.1); cfgz.want_grad = 1
dU = torch.empty_like(f)


def run_ms(iters=8):
    out = torch.zeros(lib.mmif_loss_out_doubles(B), dtype=torch.float64, device='cuda')
    ws = torch.zeros(lib.mmif_loss_workspace_bytes(B, H, W) * 4, dtype=torch.uint8, device='cuda')
    z = lambda: L.check(lib.mmif_fusion_loss_fwd(a.data_ptr(), b.data_ptr(), f.data_ptr(), B, H, W, ctypes.byref(cfgz), out.data_ptr(),
                                                  dU.data_ptr(), ws.data_ptr(), ws.numel(), st))
    for _ in range(2): z()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): z()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


os.environ.pop('MMIF_WS_GEOM', None)
print(f'model choice: {run_ms():.4f} ms')
res = []
up8 = lambda v: (v + 7) // 8 * 8
for k in (1, 2, 3, 4, 5, 6, 8):
    T = up8(-(-H // k))
    for d in (1, 2, 3, 4, 6):
        s = up8(-(-T // d))
        if s < 32: continue
        full = -(-H // T)
        for nt in (range(full, full + 1) if d == 1 else range(0, full + 1)):
            if nt * T >= H and nt != full: continue
            os.environ['MMIF_WS_GEOM'] = f'{T},{nt},{s}'
            try:
                res.append((run_ms(), T, nt, s))
            except Exception as e:
                print('fail', T, nt, s, e)
res.sort()
for ms, T, nt, s in res[:12]:
    print(f'{ms:.4f} ms  T={T} n_tall={nt} s={s}')
print('worst', res[-1])
