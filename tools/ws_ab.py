"""A/B of the warp-specialised loss+gradient kernel against the 2-CTA kernel: bit-equality of the gradient and the loss block,
and device time per launch.  MMIF_LOSS_WS is read per call by the library."""
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


shapes = [(2, 96, 160), (1, 1024, 1224), (3, 517, 1030), (2, 333, 265), (8, 256, 256), (8, 3072, 4096), (64, 3072, 4096)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in s.split('x')) for s in sys.argv[1:]]
lib = L.load()
for (B, H, W) in shapes:
    torch.manual_seed(B * H + W)
    a, b, f = (torch.rand(B, 1, H, W, device='cuda') for _ in range(3))
    st = L.stream_ptr(a.device)
    res = {}
    for mode in ('max', 'avg'):
        for ws_on in (0, 1):
            os.environ['MMIF_LOSS_WS'] = str(ws_on)
            cfgz = ML._cfg(1.0, mode, mode, 'l1' if mode == 'max' else 'l2', 'l1' if mode == 'max' else 'l2', 1.0, 0.01, 0.1)
            cfgz.want_grad = 1
            out = torch.zeros(lib.mmif_loss_out_doubles(B), dtype=torch.float64, device='cuda')
            ws = torch.zeros(lib.mmif_loss_workspace_bytes(B, H, W), dtype=torch.uint8, device='cuda')
            dU = torch.full_like(f, float('nan'))
            z = lambda: L.check(lib.mmif_fusion_loss_fwd(a.data_ptr(), b.data_ptr(), f.data_ptr(), B, H, W, ctypes.byref(cfgz), out.data_ptr(),
                                                          dU.data_ptr(), ws.data_ptr(), ws.numel(), st))
            z()
            torch.cuda.synchronize()
            res[(mode, ws_on)] = (dU.clone(), out.clone(), timeit(z, iters=5 if B * H * W > 2e8 else 20))
        d0, o0, t0 = res[(mode, 0)]
        d1, o1, t1 = res[(mode, 1)]
        nd = L.LOSS_HEAD + L.LOSS_PER_SAMPLE * B
        same_g = torch.equal(d0, d1)
        nbad = int((d0 != d1).sum().item()) if not same_g else 0
        nan1 = int(torch.isnan(d1).sum().item())
        rel = ((o0[:nd] - o1[:nd]).abs() / o0[:nd].abs().clamp_min(1e-300)).max().item()
        mp = B * H * W / 1e6
        print(f'{B}x{H}x{W} {mode}: grad bit-equal {same_g} (differ {nbad}, nan {nan1})  loss block max rel diff {rel:.2e}  '
              f'2-CTA {t0:.4f} ms ({mp / t0:.1f} Mpix/ms)  ws {t1:.4f} ms ({mp / t1:.1f} Mpix/ms)  ratio {t0 / t1:.3f}', flush=True)
