"""A/B of the backward / single-pass row-segment schemes (uniform vs tall-first two-level, fusion_loss.cu bwd_geom):
device time of the single-pass kernel for the per-rank batches of the 1/2/4/8-GPU runs of configs[4]."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, ctypes
sys.path.insert(0, %r)
import torch, mmif_b200
from mmif_b200 import _lib as L
from mmif_b200.core import loss as ML
lib = L.load()
for B in (8, 16, 32, 64):
    H, W = 3072, 4096
    a, b, f = (torch.rand(B, 1, H, W, device='cuda') for _ in range(3))
    cfg = ML._cfg(1.0, 'max', 'max', 'l1', 'l1', 1.0, 0.01, 0.1); cfg.want_grad = 1
    out = torch.empty(lib.mmif_loss_out_doubles(B), dtype=torch.float64, device='cuda')
    ws = torch.zeros(lib.mmif_loss_workspace_bytes(B, H, W), dtype=torch.uint8, device='cuda')
    dU = torch.empty_like(f); st = L.stream_ptr(a.device)
    run = lambda: L.check(lib.mmif_fusion_loss_fwd(a.data_ptr(), b.data_ptr(), f.data_ptr(), B, H, W, ctypes.byref(cfg), out.data_ptr(), dU.data_ptr(), ws.data_ptr(), ws.numel(), st))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print('%%s B=%%d: %%.3f ms  %%.1f Gpix/s  (ideal from B=64: x%%.3f)  checksum %%.9e' %% ('uniform' if os.environ.get('MMIF_UNIFORM_SEGMENTS') else 'two-level', B, ms, B * H * W / ms / 1e6, 1.0, dU.double().abs().sum().item()))
    del a, b, f, dU
''' % ROOT
for uni in (True, False):
    env = dict(os.environ)
    env['MMIF_DEBUG_GEOM'] = '1'
    if uni:
        env['MMIF_UNIFORM_SEGMENTS'] = '1'
    else:
        env.pop('MMIF_UNIFORM_SEGMENTS', None)
    subprocess.run([sys.executable, '-c', CHILD], env=env, check=False)
