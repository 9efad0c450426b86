"""Debug helper (GPU box): run the loss cases, print per-term errors and where they are."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')]
import numpy as np, torch
import cases, gates
import mmif_b200
from mmif_b200.core import loss as ML
LG = np.load(cases.HERE + '/loss_golden.npz')
names = sys.argv[1:] or cases.LOSS_CASES
for name in names:
    a, b, f = (torch.from_numpy(x).cuda() for x in cases.loss_case(name))
    f.requires_grad_(True)
    l1 = ML.SSIMLoss('ssim')(a, b, f); l2 = ML.PixelLoss('l1', 0.01)(a, b, f, mode='max'); l3 = ML.GradLoss('l1', 0.1)(a, b, f, mode='max')
    torch.cuda.synchronize()
    vals = [l1.item(), l2.item(), l3.item()]
    r32, r64 = LG[f'{name}/f32/loss'], LG[f'{name}/f64/loss']
    print(name, 'fwd', ['%.3e' % (abs(v - r) / max(abs(r), 1e-30)) for v, r in zip(vals, r64)], vals, list(r64))
    ref = LG[f'{name}/f64/grad']
    for k, t in enumerate((l1, l2, l3)):
        g = torch.autograd.grad(t, f, retain_graph=True)[0].cpu().numpy()
        frac, mx, where = gates.grad_report(g, ref[k])
        d = np.abs(g - ref[k])
        rows = np.where(d.max(axis=(0, 1, 3)) > 1e-5 * np.abs(ref[k]).max())[0]
        cols = np.where(d.max(axis=(0, 1, 2)) > 1e-5 * np.abs(ref[k]).max())[0]
        print('   grad term', k, 'bad frac %.3e max %.3e at' % (frac, mx), where, 'nan', np.isnan(g).sum(),
              'bad rows', rows[:12], '...' if len(rows) > 12 else '', 'bad cols', cols[:12], '...' if len(cols) > 12 else '')
