"""Golden vectors of the REAL reference on its own 21 sample pairs.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_samples.py

Reads data/samples/** from /root/reference with cv2 exactly as eval.py:176-181 does, stores the 42 grayscale images as
uint8 in ``samples.npz`` and, for every pair x fused-image kind (samples.KINDS), runs the reference's own
core/metric.py and core/loss.py in float32 and float64:
  * the 16-metric row of eval.py:29-75 (f32 + f64), the three marginal histograms (integer counts) and checksums of the
    two joint histograms;
  * the three loss terms of train.py:64-69 (f32 + f64), the SSIM.forward dict, and a fingerprint of the fp64 gradient
    (512 sampled elements of d(l1+l2+l3)/d imgf and of the three per-term gradients, plus their L1 norms);
  * for two small pairs: imgf = DenseFuse(seed 0)(img1, img2) from the reference's core/model.py, with the same outputs.
The GPU box never runs this script."""
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
import core.model as RMod   # noqa: E402
import samples as S         # noqa: E402
from make_golden import ref_eval_pair, loss_terms, NAMES, t   # noqa: E402


def write_samples():
    import cv2
    d = os.path.join(REF, 'data/samples')
    out, names = {}, []
    for sub, k1, k2 in (('polar', 'vis', 'po'), ('infrared', 'vis', 'ir')):
        files = sorted(os.listdir(f'{d}/{sub}/test/{k1}'))
        for n in files:
            a = cv2.imread(f'{d}/{sub}/test/{k1}/{n}', cv2.IMREAD_GRAYSCALE)
            b = cv2.imread(f'{d}/{sub}/test/{k2}/{n}', cv2.IMREAD_GRAYSCALE)
            assert a is not None and b is not None and a.shape == b.shape and a.dtype == np.uint8
            names.append(f'{sub}/{n}')
            out[f'{sub}/{n}/a'], out[f'{sub}/{n}/b'] = a, b
    out['names'] = np.array(names)
    np.savez_compressed(os.path.join(HERE, 'samples.npz'), **out)


def one_case(out, key, a, b, f, seed):
    """a, b, f: float32 (1,1,H,W) numpy on the 0..255 scale."""
    for tag, dt in (('f32', torch.float32), ('f64', torch.float64)):
        A, B, Fm = t(a, dt), t(b, dt), t(f, dt)
        with torch.no_grad():
            r = ref_eval_pair(A, B, Fm)
        out[f'{key}/{tag}/metrics'] = np.array([r[k].item() for k in NAMES], dtype=np.float64)
    A, B, Fm = t(a), t(b), t(f)
    out[f'{key}/hist'] = np.stack([torch.histc(x, 256, 0, 256).to(torch.int64).numpy() for x in (A, B, Fm)])
    for nm, x in (('af', A), ('bf', B)):
        j = np.histogram2d(x.numpy().flatten(), Fm.numpy().flatten(), 256, ((0, 256), (0, 256)))[0]
        out[f'{key}/joint_{nm}'] = S.joint_checksum(j)
    # loss convention: the same images / 255 (float32 division, data/dataset.py)
    au, bu, fu = S.unit(a), S.unit(b), S.unit(f)
    probe = S.grad_probe(fu.shape, seed)
    for tag, dt in (('f32', torch.float32), ('f64', torch.float64)):
        A, B = t(au, dt), t(bu, dt)
        Fv = t(fu, dt).clone().requires_grad_(True)
        terms = loss_terms(A, B, Fv)
        out[f'{key}/{tag}/loss'] = np.array([x.item() for x in terms], dtype=np.float64)
        grads = [torch.autograd.grad(terms[k], Fv, retain_graph=True)[0].detach().numpy().reshape(-1) for k in range(3)]
        tot = grads[0] + grads[1] + grads[2]
        out[f'{key}/{tag}/grad_probe'] = np.stack([g[probe] for g in grads] + [tot[probe]]).astype(np.float64)
        out[f'{key}/{tag}/grad_l1'] = np.array([np.abs(g.astype(np.float64)).sum() for g in grads] + [np.abs(tot.astype(np.float64)).sum()])
        out[f'{key}/{tag}/grad_max'] = np.array([np.abs(g).max() for g in grads] + [np.abs(tot).max()], dtype=np.float64)
        mod = RL.SSIM(11, 1.0)
        with torch.no_grad():
            d1, d2 = mod(A, Fv.detach()), mod(B, Fv.detach())
        out[f'{key}/{tag}/ssim_dict'] = np.array([d1['ssim'].item(), d1['cs'].item(), d1['sigma'].item(),
                                                  d2['ssim'].item(), d2['cs'].item(), d2['sigma'].item()], dtype=np.float64)


def main():
    if not os.path.exists(os.path.join(HERE, 'samples.npz')):
        write_samples()
    out = {}
    for seed, name in enumerate(S.names()):
        for kind in S.KINDS:
            a, b, f = S.case(name, kind)
            one_case(out, f'{name}/{kind}', a, b, f, seed)
            print(name, kind, out[f'{name}/{kind}/f32/loss'], flush=True)
    # imgf = DenseFuse(seed 0) output, two small pairs; the network output is stored (float32, 0..1 scale)
    for name in ('infrared/05.png', 'infrared/36.png'):
        a, b = S.pair(name)
        au, bu = S.unit(a.astype(np.float32))[None, None], S.unit(b.astype(np.float32))[None, None]
        torch.manual_seed(0)
        net = RMod.DenseFuse().eval()
        with torch.no_grad():
            res = net(t(au), t(bu))
        fu = (res['imgf'] if isinstance(res, dict) else res).numpy().astype(np.float32)
        out[f'{name}/densefuse/imgf'] = fu
        one_case(out, f'{name}/densefuse', a.astype(np.float32)[None, None], b.astype(np.float32)[None, None],
                 (fu * np.float32(255.0)).astype(np.float32), S.names().index(name))
        # the loss inputs of this case are exactly (a/255, b/255, fu*255/255): keep the convention of one_case
        print(name, 'densefuse', out[f'{name}/densefuse/f32/loss'], flush=True)
    np.savez_compressed(os.path.join(HERE, 'samples_golden.npz'), **out)
    print('written', len(out), 'arrays')


if __name__ == '__main__':
    main()
