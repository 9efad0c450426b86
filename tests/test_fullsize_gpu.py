"""Parity at BASELINE.json's full sizes (4096x3072 loss, 1224x1024 polarization suite) through size-independent
properties, where the CPU oracle on the whole input would take minutes:
  * crop consistency: the gradient is local (21x21 receptive field for SSIM, 5x5 for the Sobel term, 1x1 for
    the pixel term), so on an interior window straddling strip and row-segment boundaries the full-size GPU
    gradient must equal the fp64 ORACLE gradient of a crop with a halo, rescaled by the mean normalisers;
  * identical images: all three losses are exactly / numerically zero;
  * batch replication: loss([x, x]) == loss([x]) and grad([x, x]) == grad([x]) / 2;
  * determinism: two runs are bit-identical;
  * metric suite: qabf + nabf + labf == 1, histogram totals == pixel count, suite(pair) independent of the batch
    it travels in."""
import numpy as np
import pytest
import torch

import gates
from oracle import fusion_loss as OL

pytestmark = pytest.mark.gpu
H, W = 3072, 4096          # "4096x3072" of BASELINE configs[4]


def _mods():
    import mmif_b200  # noqa: F401
    from mmif_b200.core import loss as ML, metric as MM
    return ML, MM


def _three_grads(ML, A, B_, F_):
    l1 = ML.SSIMLoss('ssim', weight=1.0)(A, B_, F_)
    l2 = ML.PixelLoss('l1', weight=0.01)(A, B_, F_, mode='max')
    l3 = ML.GradLoss('l1', weight=0.1)(A, B_, F_, mode='max')
    gs = [torch.autograd.grad(t, F_, retain_graph=True)[0] for t in (l1, l2, l3)]
    return [l1.item(), l2.item(), l3.item()], gs


@pytest.fixture(scope='module')
def full_case():
    ML, _ = _mods()
    g = torch.Generator(device='cuda').manual_seed(99)
    a, b, f = (torch.rand(1, 1, H, W, device='cuda', generator=g) for _ in range(3))
    f = (0.5 * (a + b) + 0.2 * (f - 0.5)).contiguous()
    F_ = f.clone().requires_grad_(True)
    vals, gs = _three_grads(ML, a, b, F_)
    return a, b, f, vals, gs


@pytest.mark.parametrize('r0,c0', [(236, 60), (2290, 3960), (0, 0), (H - 120, W - 170)])
def test_full_size_gradient_equals_oracle_on_crops(full_case, r0, c0):
    """windows of 120 x 170 pixels: straddle the 104-column strips, the 256-row segments, and the image corners"""
    a, b, f, _, gs = full_case
    hh, ww, halo = 120, 170, 12
    R0, R1, C0, C1 = max(r0 - halo, 0), min(r0 + hh + halo, H), max(c0 - halo, 0), min(c0 + ww + halo, W)
    ca, cb = (t[:, :, R0:R1, C0:C1].double().cpu() for t in (a, b))
    cf = f[:, :, R0:R1, C0:C1].double().cpu().requires_grad_(True)
    terms = [OL.ssim_loss(ca, cb, cf, 'ssim', 1.0, False, 1.0), OL.pixel_loss(ca, cb, cf, 'l1', 0.01, 'max'),
             OL.grad_loss(ca, cb, cf, 'l1', 0.1, 'max')]
    ch, cw = R1 - R0, C1 - C0
    scale = [(ch - 10) * (cw - 10) / ((H - 10) * (W - 10)), ch * cw / (H * W), ch * cw / (H * W)]   # mean normalisers
    # interior of the crop whose receptive field lies inside the crop, except where the crop edge IS the image edge
    i0, i1 = (0 if R0 == 0 else r0 - R0), (ch if R1 == H else r0 - R0 + hh)
    j0, j1 = (0 if C0 == 0 else c0 - C0), (cw if C1 == W else c0 - C0 + ww)
    for k, nm in enumerate(('ssim', 'pixel', 'grad')):
        g64, = torch.autograd.grad(terms[k], cf, retain_graph=True)
        ref = g64[0, 0, i0:i1, j0:j1].numpy() * scale[k]
        got = gs[k][0, 0, R0 + i0:R0 + i1, C0 + j0:C0 + j1].cpu().numpy()
        frac, mx, where = gates.grad_report(got, ref)
        if nm == 'ssim':
            assert mx <= 2e-5, f'{nm}: max-norm err {mx:.3e} at {where}'
        else:
            assert frac <= 1e-4, f'{nm}: {frac:.2e} of elements differ (sign ties), max {mx:.3e} at {where}'


def test_full_size_identity_replication_determinism(full_case):
    ML, _ = _mods()
    a, b, f, vals, gs = full_case
    F2 = f.clone().requires_grad_(True)
    vals2, gs2 = _three_grads(ML, a, b, F2)
    assert vals == vals2 and all(torch.equal(x, y) for x, y in zip(gs, gs2))                  # deterministic
    A2, B2 = torch.cat([a, a]), torch.cat([b, b])
    F3 = torch.cat([f, f]).requires_grad_(True)
    vals3, gs3 = _three_grads(ML, A2, B2, F3)
    for k in range(3):
        assert abs(vals3[k] - vals[k]) <= 1e-6 * abs(vals[k])
        # the segment geometry (hence the tile shifts and the rounding) may change with the batch size: max-norm gate
        err = (gs3[k][0] - gs[k][0] * 0.5).abs().max().item()
        assert err <= 1e-5 * 0.5 * gs[k][0].abs().max().item() and torch.equal(gs3[k][0], gs3[k][1])
    same, _ = _three_grads(ML, a, a, a.clone().requires_grad_(True))
    assert abs(same[0]) <= 1e-6 and same[1] == 0.0 and same[2] == 0.0                        # ssim(x,x)=1, |x-max(x,x)|=0
    tot = float(np.sum(vals))
    assert np.isfinite(tot) and all(torch.isfinite(g).all().item() for g in gs)


def test_polar_suite_invariants_at_full_batch():
    """BASELINE configs[3]: 32 pairs of 1224x1024."""
    _, MM = _mods()
    g = torch.Generator(device='cuda').manual_seed(5)
    n, h, w = 32, 1024, 1224
    a = torch.randint(0, 256, (n, 1, h, w), device='cuda', generator=g).float()
    b = torch.randint(0, 256, (n, 1, h, w), device='cuda', generator=g).float()
    f = torch.floor((a + b) / 2)
    rows = MM.eval_metrics_batch(a, b, f)
    assert torch.equal(rows, MM.eval_metrics_batch(a, b, f))
    r = rows.cpu().numpy()
    assert np.isfinite(r).all()
    np.testing.assert_allclose(r[:, 10] + r[:, 11] + r[:, 12], 1.0, rtol=0, atol=2e-6)        # qabf + nabf + labf = 1
    one = MM.eval_metrics_batch(a[7:8], b[7:8], f[7:8])[0].cpu().numpy()
    np.testing.assert_allclose(r[7], one, rtol=2e-6, atol=1e-12)      # (segment geometry depends on the batch size)
    counts, _ = MM.hist_raw(a, b, f)
    c = counts.to(torch.int64)
    assert (c[:, 0:256].sum(1) == h * w).all() and (c[:, 768:768 + 65536].sum(1) == h * w).all()
    assert torch.equal(c[:, 768:768 + 65536].view(n, 256, 256).sum(1), c[:, 512:768])       # f marginal = column sums of joint(a,f)


def test_full_size_whole_image_vs_fp64_oracle_on_the_gpu():
    """The WHOLE 4096x3072 gradient and the three losses against the oracle evaluated in float64 with the same torch
    ops on the GPU (the oracle is dtype/device generic; its fp64 CUDA graph is an independent implementation: cuDNN /
    ATen kernels).  Gate: every element within 1e-5 max|g64| except L1 sign ties (<= 1e-4 of the elements), losses
    within 1e-5 relative."""
    ML, _ = _mods()
    g = torch.Generator(device='cuda').manual_seed(5)
    a, b, f = (torch.rand(1, 1, H, W, device='cuda', generator=g) for _ in range(3))
    F_ = f.clone().requires_grad_(True)
    l1 = ML.SSIMLoss('ssim', weight=1.0)(a, b, F_)
    l2 = ML.PixelLoss('l1', weight=0.01)(a, b, F_, mode='max')
    l3 = ML.GradLoss('l1', weight=0.1)(a, b, F_, mode='max')
    (l1 + l2 + l3).backward()
    keep = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        l64, g64 = OL.train_objective_grad(a.double(), b.double(), f.double())
    finally:
        torch.backends.cudnn.allow_tf32 = keep
    for nm, new, ref in zip(('ssim', 'pixel', 'grad'), (l1, l2, l3), l64):
        assert abs(new.item() - ref.item()) <= 1e-5 * abs(ref.item()), (nm, new.item(), ref.item())
    d = (F_.grad.double() - g64).abs()
    scale = g64.abs().max()
    frac_bad = (d > 1e-5 * scale).double().mean().item()
    assert frac_bad <= 1e-4, (frac_bad, (d.max() / scale).item())
    # away from the sign ties the agreement is ~5e-7 (measured): the median error must be far below the gate
    assert (d.median() / scale).item() < 1e-6
