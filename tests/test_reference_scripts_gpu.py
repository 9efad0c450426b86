"""The reference's UNMODIFIED train.py, test.py and eval.py (byte-compiled by oracle/build_ref.py into oracle/_ref/scripts,
which travels to the GPU box like a built .so) run end to end with `dropin/` first on sys.path — the zero-change drop-in of
INTEGRATION.md section 1 — on a small dataset made from the reference's own sample pairs:

    train.py  (train.py:37-133, 302-317: DeepFuse, SSIMLoss + PixelLoss + GradLoss, backward, clip, Adam; 1 epoch)
 -> test.py   (test.py:31-69: fused images + calc_ssim of core.metric)
 -> eval.py   (eval.py:29-75, 150-361: the 16 metrics over the fused images, written through openpyxl)

and, for the loss, the same training run with the reference's OWN core/loss.py (torch CUDA eager) as the comparison.
natsort / openpyxl / patchify / thop are not installed in this image: tests/stubs provides minimal stand-ins."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import samples as S

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRAIN = ['infrared/05.png', 'infrared/36.png', 'infrared/175.png', 'infrared/037.png', 'infrared/049.png', 'infrared/100.png',
         'infrared/108.png', 'infrared/17.png', 'infrared/18.png', 'infrared/00537D.png']
TEST = ['infrared/00556D.png', 'infrared/21.png', 'infrared/00633D.png']


def _stage(tmp, with_reference_loss):
    import cv2
    from oracle import build_ref
    repo = os.path.join(tmp, 'repo')
    build_ref.stage_scripts(repo, with_reference_loss=with_reference_loss)
    for split, names in (('train', TRAIN), ('test', TEST)):
        for sub in ('vis', 'ir'):
            os.makedirs(os.path.join(tmp, 'datasets', 'roadscene', split, sub), exist_ok=True)
        for k, n in enumerate(names):
            a, b = S.pair(n)
            cv2.imwrite(os.path.join(tmp, 'datasets', 'roadscene', split, 'vis', f'{k + 1}.png'), a)
            cv2.imwrite(os.path.join(tmp, 'datasets', 'roadscene', split, 'ir', f'{k + 1}.png'), b)
    return repo


def _run(repo, script, args, dropin, counts=None):
    env = dict(os.environ)
    path = [os.path.join(ROOT, 'tests', 'stubs'), repo]
    if dropin:
        path.insert(0, os.path.join(ROOT, 'dropin'))
    env['PYTHONPATH'] = os.pathsep.join(path)
    env['PYTHONDONTWRITEBYTECODE'] = '1'
    if counts:
        env['MMIF_COUNTS_FILE'] = counts
    out = subprocess.run([sys.executable, os.path.join(repo, script + '.pyc')] + args, capture_output=True, text=True, env=env,
                         cwd=repo, timeout=900)
    assert out.returncode == 0, f'{script} failed:\n{out.stdout[-3000:]}\n{out.stderr[-3000:]}'
    return out.stdout + out.stderr


def _only_ckpt(tmp):
    d = os.path.join(tmp, 'checkpoints')
    names = sorted(os.listdir(d))
    assert len(names) == 1, names
    return names[0], os.path.join(d, names[0])


def _epoch_losses(ckpt_dir):
    log = open(os.path.join(ckpt_dir, 'train.log')).read()
    m = re.search(r'epoch: 01, train loss: ([0-9.]+), valid loss: ([0-9.]+)', log)
    assert m, log[-2000:]
    return float(m.group(1)), float(m.group(2))


def test_unmodified_train_test_eval_with_the_dropin(tmp_path):
    from oracle import build_ref
    if not build_ref.scripts_available():
        pytest.skip('oracle/_ref/scripts not built (run oracle/build_ref.py where /root/reference exists)')
    import cv2
    import torch
    tmp = str(tmp_path / 'dropin')
    repo = _stage(tmp, with_reference_loss=False)
    assert not os.path.exists(os.path.join(repo, 'core', 'loss.pyc'))       # core.loss / core.metric can only come from dropin/
    counts = os.path.join(tmp, 'counts_train.json')
    _run(repo, 'train', ['--epoch', '1', '--bs', '4', '--use_patches', '', '--data', 'roadscene'], dropin=True, counts=counts)
    ckpt, ckpt_dir = _only_ckpt(tmp)
    tl, vl = _epoch_losses(ckpt_dir)
    assert np.isfinite(tl) and np.isfinite(vl) and 0.0 < tl < 3.0
    c = json.load(open(counts))
    # every training iteration = ONE single-pass loss+gradient launch + the in-place rescale; validation = forward-only kernel
    assert c['loss_single_pass'] >= 2 and c['rescale'] == c['loss_single_pass'] and c['loss_bwd'] == 0 and c['loss_fwd'] >= 1, c
    assert c['tmap_fail'] == 0, c
    assert os.path.isfile(os.path.join(ckpt_dir, 'epoch_best.pth'))

    # ---- the same run with the reference's own core/loss.py (torch CUDA eager): same seed, same data -> same losses
    tmp_ref = str(tmp_path / 'reference')
    repo_ref = _stage(tmp_ref, with_reference_loss=True)
    _run(repo_ref, 'train', ['--epoch', '1', '--bs', '4', '--use_patches', '', '--data', 'roadscene'], dropin=False)
    tl_ref, vl_ref = _epoch_losses(_only_ckpt(tmp_ref)[1])
    assert abs(tl - tl_ref) <= 2e-3 * abs(tl_ref) + 1e-4, (tl, tl_ref)      # 4 decimals are logged; cuDNN TF32 in the network
    assert abs(vl - vl_ref) <= 5e-3 * abs(vl_ref) + 1e-4, (vl, vl_ref)

    # ---- test.py: fused images + the SSIM it prints (core.metric.calc_ssim on CUDA tensors)
    counts = os.path.join(tmp, 'counts_test.json')
    out = _run(repo, 'test', ['--data', 'roadscene', '--ckpt', ckpt], dropin=True, counts=counts)
    ssims = [float(v) for v in re.findall(r'iter: \d+, ssim: ([0-9.]+)', out)]
    assert len(ssims) == len(TEST) and all(0.0 < v <= 1.0 for v in ssims), out[-2000:]
    fused_dir = os.path.join(ckpt_dir, 'roadscene')
    assert sorted(os.listdir(fused_dir)) == ['01.bmp', '02.bmp', '03.bmp']
    assert json.load(open(counts))['moment_fwd'] >= len(TEST)
    # sanity only: the printed SSIM is of the network's float output, the file holds its clip(0,1) 8-bit version (an untrained
    # DeepFuse leaves [0,1] in places); exact calc_ssim parity is the business of test_metric_gpu / test_samples_gpu
    from oracle import fusion_metric as OM
    a, b = (torch.from_numpy(x.astype(np.float32))[None, None] for x in S.pair(TEST[0]))
    f8 = torch.from_numpy(cv2.imread(os.path.join(fused_dir, '01.bmp'), cv2.IMREAD_GRAYSCALE).astype(np.float32))[None, None]
    ref = 0.5 * (OM.ssim(a / 255.0, f8 / 255.0, data_range=1.0) + OM.ssim(b / 255.0, f8 / 255.0, data_range=1.0)).item()
    assert abs(ssims[0] - ref) <= 0.05, (ssims[0], ref)

    # ---- eval.py over the fused images test.py wrote: the sheet equals the oracle's rows on the same files
    counts = os.path.join(tmp, 'counts_eval.json')
    _run(repo, 'eval', ['--data', 'roadscene', '--ckpt', ckpt], dropin=True, counts=counts)
    sheet = json.load(open(os.path.join(ckpt_dir, 'metrics_roadscene_DeepFuse.xlsx')))['DeepFuse']
    assert [sheet[f'{c}1'] for c in 'BCDEFGHIJKLMNOPQ'] == ['SD', 'AG', 'SF', 'MSE', 'PSNR', 'CC', 'SCD', 'EN', 'CE', 'MI', 'Qabf', 'Nabf',
                                                              'Labf', 'SSIM', 'MSSSIM', 'VIFF']
    assert [sheet['A2'], sheet['A3'], sheet['A4']] == ['mean', 'std', '1.png']
    import gates
    for k, n in enumerate(TEST):
        a, b = (torch.from_numpy(x.astype(np.float32))[None, None] for x in S.pair(n))
        f = torch.from_numpy(cv2.imread(os.path.join(fused_dir, f'{k + 1:0>2}.bmp'), cv2.IMREAD_GRAYSCALE).astype(np.float32))[None, None]
        r32, r64 = OM.eval_pair(a, b, f), OM.eval_pair(a.double(), b.double(), f.double())
        for col, nm in zip('BCDEFGHIJKLMNOPQ', OM.METRIC_NAMES):
            allow = None
            if nm in ('nabf', 'labf'):
                allow = lambda a=a, b=b, f=f: gates.qabf_tie_allowance(a.numpy(), b.numpy(), f.numpy())
            if nm == 'viff':
                allow = lambda a=a, b=b, f=f: gates.viff_tie_allowance(a.numpy(), b.numpy(), f.numpy())
            gates.assert_scalar(f'eval.py sheet {n}/{nm}', sheet[f'{col}{k + 4}'], r32[nm], r64[nm], allowance=allow)
    c = json.load(open(counts))
    assert c['metric'] > 0 and c['moment_fwd'] > 0, c
