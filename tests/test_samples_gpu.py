"""GPU parity on the reference's OWN 21 sample pairs (data/samples/**, committed as uint8 in tests/golden/samples.npz) x three
synthetic fused images each (max, rounded average, average + noise) + two DenseFuse(seed 0) outputs,
against golden vectors the REAL reference produced (tests/golden/make_golden_samples.py): the full 1024x1224 polar pairs
(where the reference's own fp32 noise is largest) and the five infrared pairs whose width is not a multiple of 4 (no TMA:
the plain-load ring) included.  Metrics through the batched suite entry (CUDA tensors) and through the per-function
drop-ins as eval.py calls them (CPU tensors); histograms bit-exact; losses + gradient through the drop-in modules."""
import numpy as np
import pytest
import torch

import gates
import samples as S
from oracle import fusion_metric as OM

pytestmark = pytest.mark.gpu
SG = np.load(S.HERE + '/samples_golden.npz')
NAMES = S.names()
CASES = [(n, k) for n in NAMES for k in S.KINDS]


def _mods():
    import mmif_b200  # noqa: F401
    from mmif_b200 import _lib as L
    from mmif_b200.core import loss as ML, metric as MM
    return L, ML, MM


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x)).float()


def _allowance(name, kind, metric):
    """Tie allowances of the two metric families with a hard per-pixel selection (evaluated only if the plain gate fails)."""
    if metric in ('nabf', 'labf'):
        return lambda: gates.qabf_tie_allowance(*S.case(name, kind))
    if metric == 'viff':
        return lambda: gates.viff_tie_allowance(*S.case(name, kind))
    return None


@pytest.mark.parametrize('name,kind', CASES)
def test_metric_row_cuda_inputs(name, kind):
    _, _, MM = _mods()
    a, b, f = (T(x).cuda() for x in S.case(name, kind))
    row = MM.eval_metrics_batch(a, b, f)[0].cpu().numpy()
    r32, r64 = SG[f'{name}/{kind}/f32/metrics'], SG[f'{name}/{kind}/f64/metrics']
    for k, nm in enumerate(OM.METRIC_NAMES):
        gates.assert_scalar(f'{name}/{kind}/{nm}', row[k], r32[k], r64[k], allowance=_allowance(name, kind, nm))


@pytest.mark.parametrize('name,kind', CASES)
def test_histograms_bit_exact(name, kind):
    _, _, MM = _mods()
    a, b, f = (T(x).cuda() for x in S.case(name, kind))
    ha, hb, hf, jaf, jbf = MM.histograms(a, b, f)
    assert np.array_equal(np.stack([ha.numpy(), hb.numpy(), hf.numpy()]), SG[f'{name}/{kind}/hist'])
    assert np.array_equal(S.joint_checksum(jaf.numpy()), SG[f'{name}/{kind}/joint_af'])
    assert np.array_equal(S.joint_checksum(jbf.numpy()), SG[f'{name}/{kind}/joint_bf'])
    assert int(jaf.sum()) <= a.numel() and int(ha.sum()) == a.numel()      # 8-bit sources: nothing dropped from the marginals


@pytest.mark.parametrize('name', NAMES)
def test_eval_py_call_pattern_cpu_inputs(name):
    """The calls eval.py:29-75 makes, on CPU tensors as eval.py:189-200 leaves them (the drop-ins upload them)."""
    _, _, MM = _mods()
    kind = S.KINDS[NAMES.index(name) % 3]
    a, b, f = (T(x) for x in S.case(name, kind))
    m = (MM.calc_mse(a, f) + MM.calc_mse(b, f)) * 0.5
    q, n, l = MM.calc_Qabf(a, b, f, L=1.5, full=True)
    vals = [MM.calc_std(f), MM.calc_ag(f), MM.calc_sf(f), m, MM.calc_psnr(m), (MM.calc_cc(a, f) + MM.calc_cc(b, f)) * 0.5,
            MM.calc_scd(a, b, f), MM.calc_entropy(f), MM.calc_cross_ent(a, f) + MM.calc_cross_ent(b, f),
            MM.calc_mul_info(a, f, normalized=True) + MM.calc_mul_info(b, f, normalized=True), q, n, l,
            (MM.calc_ssim(a, f) + MM.calc_ssim(b, f)) * 0.5, (MM.calc_msssim(a, f) + MM.calc_msssim(b, f)) * 0.5,
            MM.calc_viff(a, b, f, simple=False)]
    assert all(isinstance(v, torch.Tensor) and v.dim() == 0 and not v.is_cuda for v in vals)
    assert vals[9].dtype == torch.float64
    r32, r64 = SG[f'{name}/{kind}/f32/metrics'], SG[f'{name}/{kind}/f64/metrics']
    for k, nm in enumerate(OM.METRIC_NAMES):
        gates.assert_scalar(f'{name}/{kind}/{nm} (drop-in, cpu tensors)', vals[k].item(), r32[k], r64[k],
                            allowance=_allowance(name, kind, nm))


def _loss_case(name, kind):
    if kind == 'densefuse':
        a, b = (x.astype(np.float32)[None, None] for x in S.pair(name))
        f = (SG[f'{name}/densefuse/imgf'] * np.float32(255.0)).astype(np.float32)
    else:
        a, b, f = S.case(name, kind)
    return S.unit(a), S.unit(b), S.unit(f)


@pytest.mark.parametrize('name,kind', CASES + [('infrared/05.png', 'densefuse'), ('infrared/36.png', 'densefuse')])
def test_loss_and_gradient_through_the_modules(name, kind):
    """train.py:64-71 on real images: the three losses against the reference's fp32 / fp64 values; d(total)/d imgf at the
    512 probed elements against the reference's fp64 gradient (elements an L1 sign tie can touch get the tie allowance)."""
    L, ML, _ = _mods()
    a, b, f = _loss_case(name, kind)
    A, B_, F_ = T(a).cuda(), T(b).cuda(), T(f).cuda().requires_grad_(True)
    c0 = L.launch_counts()
    l1 = ML.SSIMLoss('ssim', weight=1.0)(A, B_, F_)
    l2 = ML.PixelLoss('l1', weight=0.01)(A, B_, F_, mode='max')
    l3 = ML.GradLoss('l1', weight=0.1)(A, B_, F_, mode='max')
    (l1 + l2 + l3).backward()
    c1 = L.launch_counts()
    assert c1['loss_single_pass'] - c0['loss_single_pass'] == 1 and c1['loss_fwd'] == c0['loss_fwd']
    r32, r64 = SG[f'{name}/{kind}/f32/loss'], SG[f'{name}/{kind}/f64/loss']
    for k, (nm, v) in enumerate(zip(('ssim', 'pixel', 'grad'), (l1, l2, l3))):
        gates.assert_scalar(f'{name}/{kind}/{nm}', v.item(), r32[k], r64[k])
    got = F_.grad.cpu().numpy().reshape(-1).astype(np.float64)
    assert np.isfinite(got).all()
    probe = S.grad_probe(f.shape, NAMES.index(name))
    p64, p32 = SG[f'{name}/{kind}/f64/grad_probe'], SG[f'{name}/{kind}/f32/grad_probe']
    gmax = SG[f'{name}/{kind}/f64/grad_max'][3]
    pix_mask, sob_mask, pix_zero = gates.l1_tie_masks(a, b, f)
    tied = (pix_mask | sob_mask).reshape(-1)[probe]
    diff = np.abs(got[probe] - p64[3])
    ref_diff = np.abs(p32[3] - p64[3])
    # untied elements: 1e-5 of max|g|, or the reference's own fp32 error at the untied probes where that is larger
    rtol = max(gates.RTOL, (ref_diff * ~tied).max() / gmax)
    assert (diff * ~tied).max() <= rtol * gmax, f'{name}/{kind}: untied probe error {(diff * ~tied).max() / gmax:.3e} of max|g| (reference fp32: {(ref_diff * ~tied).max() / gmax:.3e})'
    flip = (64 * 0.1 + 2 * 0.01) / f.size
    assert (diff * tied).max() <= rtol * gmax + flip * (1 + 1e-5), f'{name}/{kind}: a tied probe moved by {(diff * tied).max():.3e}'
    if kind == 'max':        # imgf == max(img1, img2) everywhere: the pixel term vanishes and its gradient is exactly zero
        assert l2.item() == 0.0 and pix_zero.all()
