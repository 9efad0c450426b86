"""Golden vectors for SSIM(win_size) / MS_SSIM(win_size) with the 9/7/5/3-tap windows, from the REAL reference.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_windows.py

Imports ``core/loss.py`` from ``/root/reference`` (read-only, never copied) in the build container and stores, for the
seeded inputs below, the per-sample dict of ``SSIM(win_size=k)(a, f)`` (float32 = the reference's arithmetic, float64 =
inputs cast to double) with and without padding, and ``MS_SSIM(win_size=k)(a, f)`` — in ``ssim_windows_golden.npz``
together with the inputs.  The GPU box only reads the committed file.
"""
import os
import sys

sys.dont_write_bytecode = True
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '/root/reference')
sys.path.insert(0, HERE)
import core.loss as RL      # noqa: E402  (the reference)
import cases                # noqa: E402


def main():
    out = {}
    a, _, f = (torch.from_numpy(np.ascontiguousarray(x)) for x in cases.loss_case('rand_3x64x96'))
    for win in (9, 7, 5, 3):
        for pad in (False, True):
            for tag, dt in (('f32', torch.float32), ('f64', torch.float64)):
                mod = RL.SSIM(win, 1.0, pad).to(dt)
                d = mod(a.to(dt), f.to(dt))
                out[f'ssim/win{win}/pad{int(pad)}/{tag}'] = np.stack([d[k].detach().numpy().astype(np.float64) for k in ('ssim', 'cs', 'sigma')])
    g = torch.Generator().manual_seed(77)
    x, y = (torch.rand(2, 1, 208, 240, generator=g) for _ in range(2))
    y = (0.6 * x + 0.4 * y).contiguous()
    out['ms/x'], out['ms/y'] = x.numpy(), y.numpy()
    for win, pad in ((7, False), (5, True)):
        for tag, dt in (('f32', torch.float32), ('f64', torch.float64)):
            mod = RL.MS_SSIM(win, 1.0, pad).to(dt)
            out[f'msssim/win{win}/pad{int(pad)}/{tag}'] = mod(x.to(dt), y.to(dt)).detach().numpy().astype(np.float64)
    np.savez_compressed(os.path.join(HERE, 'ssim_windows_golden.npz'), **out)
    print('wrote', len(out), 'arrays')


if __name__ == '__main__':
    main()
