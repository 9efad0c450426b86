"""Generate the golden vectors by running the REAL reference in the build container.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports ``core/loss.py`` and ``core/metric.py`` from ``/root/reference`` (read-only,
never copied), runs them on the seeded inputs of ``cases.py`` in float32 (the
reference's arithmetic) and float64 (inputs cast to double; every reference
function is dtype-generic), and stores the outputs in ``loss_golden.npz`` /
``metric_golden.npz``.  Also writes ``crops.npz`` (small uint8 crops of two of the
reference's sample image pairs) the first time it runs.  The GPU box never runs
this script: it only reads the committed ``.npz`` files.
"""
import os
import sys

sys.dont_write_bytecode = True
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
sys.path.insert(0, REF)
sys.path.insert(0, HERE)

import core.loss as RL      # noqa: E402  (the reference)
import core.metric as RM    # noqa: E402


def make_crops():
    import cv2
    d = os.path.join(REF, 'data/samples')
    vis = cv2.imread(os.path.join(d, 'polar/test/vis/1.jpg'), cv2.IMREAD_GRAYSCALE)
    po = cv2.imread(os.path.join(d, 'polar/test/po/1.jpg'), cv2.IMREAD_GRAYSCALE)
    iv = cv2.imread(os.path.join(d, 'infrared/test/vis/17.png'), cv2.IMREAD_GRAYSCALE)
    ii = cv2.imread(os.path.join(d, 'infrared/test/ir/17.png'), cv2.IMREAD_GRAYSCALE)
    np.savez_compressed(os.path.join(HERE, 'crops.npz'),
                        polar_vis=vis[400:512, 500:660], polar_po=po[400:512, 500:660],
                        ir_vis=iv[60:165, 100:241], ir_ir=ii[60:165, 100:241])


def t(x, dt=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dt)


def ref_eval_pair(a, b, f):
    """eval.py:29-75 composed from the reference's own functions."""
    mse = (RM.calc_mse(a, f) + RM.calc_mse(b, f)) * 0.5
    q, n, l = RM.calc_Qabf(a, b, f, L=1.5, full=True)
    r = {'sd': RM.calc_std(f), 'ag': RM.calc_ag(f), 'sf': RM.calc_sf(f), 'mse': mse,
         'psnr': RM.calc_psnr(mse), 'cc': (RM.calc_cc(a, f) + RM.calc_cc(b, f)) * 0.5,
         'scd': RM.calc_scd(a, b, f), 'en': RM.calc_entropy(f),
         'ce': RM.calc_cross_ent(a, f) + RM.calc_cross_ent(b, f),
         'mi': RM.calc_mul_info(a, f, normalized=True) + RM.calc_mul_info(b, f, normalized=True),
         'qabf': q, 'nabf': n, 'labf': l,
         'ssim': (RM.calc_ssim(a, f) + RM.calc_ssim(b, f)) * 0.5,
         'msssim': (RM.calc_msssim(a, f) + RM.calc_msssim(b, f)) * 0.5,
         'viff': RM.calc_viff(a, b, f, simple=False)}
    return r


NAMES = ('sd', 'ag', 'sf', 'mse', 'psnr', 'cc', 'scd', 'en', 'ce', 'mi', 'qabf', 'nabf', 'labf',
         'ssim', 'msssim', 'viff')


def loss_terms(a, b, f):
    l1 = RL.SSIMLoss('ssim', weight=1.0)(a, b, f)
    l2 = RL.PixelLoss('l1', weight=0.01)(a, b, f, mode='max')
    l3 = RL.GradLoss('l1', weight=0.1)(a, b, f, mode='max')
    return l1, l2, l3


def main():
    if not os.path.exists(os.path.join(HERE, 'crops.npz')):
        make_crops()
    import cases
    out = {}
    for name in cases.LOSS_CASES:
        a, b, f = (t(x) for x in cases.loss_case(name))
        for tag, dt in (('f32', torch.float32), ('f64', torch.float64)):
            A, B = a.to(dt), b.to(dt)
            Fv = f.to(dt).clone().requires_grad_(True)
            terms = loss_terms(A, B, Fv)
            grads = []
            for k in range(3):
                g, = torch.autograd.grad(terms[k], Fv, retain_graph=True)
                grads.append(g.detach().numpy())
            out[f'{name}/{tag}/loss'] = np.array([x.item() for x in terms], dtype=np.float64)
            if tag == 'f64':
                out[f'{name}/{tag}/grad'] = np.stack(grads)  # (3,B,1,H,W) per-term d/d imgf
            else:
                out[f'{name}/{tag}/grad_total'] = (grads[0] + grads[1] + grads[2])
            mod = RL.SSIM(11, 1.0)
            d1, d2 = mod(A, Fv.detach()), mod(B, Fv.detach())
            out[f'{name}/{tag}/ssim_dict'] = np.stack(
                [d1['ssim'].numpy(), d1['cs'].numpy(), d1['sigma'].numpy(),
                 d2['ssim'].numpy(), d2['cs'].numpy(), d2['sigma'].numpy()]).astype(np.float64)
            # secondary modes (oracle pin only)
            sec = [RL.PixelLoss('l1', 0.01)(A, B, Fv.detach(), mode='avg').item(),
                   RL.PixelLoss('l2', 0.01)(A, B, Fv.detach(), mode='max').item(),
                   RL.GradLoss('l1', 0.1)(A, B, Fv.detach(), mode='avg').item(),
                   RL.GradLoss('l2', 0.1)(A, B, Fv.detach(), mode='max').item(),
                   RL.SSIMLoss('w-ssim')(A, B, Fv.detach()).item(),
                   RL.TVLoss('l1')(Fv.detach() - A).item()]
            out[f'{name}/{tag}/secondary'] = np.array(sec, dtype=np.float64)
    # reference smoke main of loss.py:388-423 (torch generator, 'avg' modes)
    torch.manual_seed(0)
    x1, x2, y = torch.rand(2, 1, 256, 256), torch.rand(2, 1, 256, 256), torch.rand(2, 1, 256, 256)
    out['smoke_main/inputs_checksum'] = np.array([x1.double().sum().item(), x2.double().sum().item(),
                                                  y.double().sum().item()])
    out['smoke_main/loss'] = np.array([RL.SSIMLoss('ssim', weight=1.0)(x1, x2, y).item(),
                                       RL.PixelLoss('l1', weight=0.01)(x1, x2, y).item(),
                                       RL.GradLoss('l1', weight=0.1)(x1, x2, y).item(),
                                       RL.TVLoss('l1', weight=1.0)(y - x1).item()])
    np.savez_compressed(os.path.join(HERE, 'loss_golden.npz'), **out)

    out = {}
    for name in cases.METRIC_CASES:
        a, b, f = (t(x) for x in cases.metric_case(name))
        for tag, dt in (('f32', torch.float32), ('f64', torch.float64)):
            r = ref_eval_pair(a.to(dt), b.to(dt), f.to(dt))
            out[f'{name}/{tag}/metrics'] = np.array([r[k].item() for k in NAMES], dtype=np.float64)
            extra = [RM.calc_mean(f.to(dt)).item(),
                     RM.calc_Nabf(a.to(dt), b.to(dt), f.to(dt), modified=False).item(),
                     RM.calc_viff(a.to(dt), b.to(dt), f.to(dt), simple=True).item(),
                     RM.calc_mul_info(a.to(dt), f.to(dt)).item(),
                     RM.calc_ssim(a.to(dt) / 255.0, f.to(dt) / 255.0, data_range=1.0).item(),   # test.py:51-52 convention
                     RM.calc_psnr(RM.calc_mse(a.to(dt), f.to(dt)), root=True).item()]
            out[f'{name}/{tag}/extra'] = np.array(extra, dtype=np.float64)
        out[f'{name}/hist_a'] = torch.histc(a, 256, 0, 256).to(torch.int64).numpy()
        out[f'{name}/hist_b'] = torch.histc(b, 256, 0, 256).to(torch.int64).numpy()
        out[f'{name}/hist_f'] = torch.histc(f, 256, 0, 256).to(torch.int64).numpy()
        ja = np.histogram2d(a.numpy().flatten(), f.numpy().flatten(), 256, ((0, 256), (0, 256)))[0]
        jb = np.histogram2d(b.numpy().flatten(), f.numpy().flatten(), 256, ((0, 256), (0, 256)))[0]
        nz = np.nonzero(ja)
        out[f'{name}/joint_af_idx'] = (nz[0] * 256 + nz[1]).astype(np.int32)
        out[f'{name}/joint_af_cnt'] = ja[nz].astype(np.int64)
        nz = np.nonzero(jb)
        out[f'{name}/joint_bf_idx'] = (nz[0] * 256 + nz[1]).astype(np.int32)
        out[f'{name}/joint_bf_cnt'] = jb[nz].astype(np.int64)
    v, w = cases.hist_edge_vector()
    out['hist_edge/hist_v'] = torch.histc(t(v), 256, 0, 256).to(torch.int64).numpy()
    out['hist_edge/hist_w'] = torch.histc(t(w), 256, 0, 256).to(torch.int64).numpy()
    j = np.histogram2d(v.flatten(), w.flatten(), 256, ((0, 256), (0, 256)))[0]
    nz = np.nonzero(j)
    out['hist_edge/joint_idx'] = (nz[0] * 256 + nz[1]).astype(np.int32)
    out['hist_edge/joint_cnt'] = j[nz].astype(np.int64)
    # reference smoke main of metric.py:494-551
    torch.manual_seed(0)
    x1, x2, y = (torch.rand(1, 1, 256, 256) * 255.0 for _ in range(3))
    r = ref_eval_pair(x1, x2, y)
    out['smoke_main/metrics'] = np.array([r[k].item() for k in NAMES], dtype=np.float64)
    out['smoke_main/inputs_checksum'] = np.array([x1.double().sum().item(), x2.double().sum().item(),
                                                  y.double().sum().item()])
    np.savez_compressed(os.path.join(HERE, 'metric_golden.npz'), **out)
    print('golden written:', [f for f in os.listdir(HERE) if f.endswith('.npz')])


if __name__ == '__main__':
    main()
