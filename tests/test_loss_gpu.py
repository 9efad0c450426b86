"""GPU parity of the fused fusion objective (forward + backward) against the oracle and the
golden vectors of the real reference.  Everything goes through the drop-in modules, i.e. through
the C ABI of libmmif_b200.so."""
import numpy as np
import pytest
import torch

import cases
import gates
from oracle import fusion_loss as OL

pytestmark = pytest.mark.gpu
LG = np.load(cases.HERE + '/loss_golden.npz')


WG = np.load(cases.HERE + '/ssim_windows_golden.npz')     # SSIM / MS_SSIM with 9/7/5/3-tap windows, from the real reference


def _mods():
    import mmif_b200  # noqa: F401
    from mmif_b200.core import loss as ML
    return ML


def T(x, dt=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dt)


def run_new(a, b, f, need_grad=True, pixel=('l1', 'max'), grad=('l1', 'max'), w=(1.0, 0.01, 0.1)):
    ML = _mods()
    A, B_, F_ = a.cuda(), b.cuda(), f.cuda().requires_grad_(need_grad)
    l1 = ML.SSIMLoss('ssim', weight=w[0])(A, B_, F_)
    l2 = ML.PixelLoss(pixel[0], weight=w[1])(A, B_, F_, mode=pixel[1])
    l3 = ML.GradLoss(grad[0], weight=w[2])(A, B_, F_, mode=grad[1])
    grads = None
    if need_grad:
        grads = [torch.autograd.grad(t, F_, retain_graph=True)[0].cpu().numpy() for t in (l1, l2, l3)]
    return [l1.item(), l2.item(), l3.item()], grads


@pytest.mark.parametrize('name', cases.LOSS_CASES)
def test_loss_forward_vs_golden(name):
    a, b, f = (T(x) for x in cases.loss_case(name))
    vals, _ = run_new(a, b, f, need_grad=False)
    for k, nm in enumerate(('ssim', 'pixel', 'grad')):
        gates.assert_scalar(f'{name}/{nm}', vals[k], LG[f'{name}/f32/loss'][k], LG[f'{name}/f64/loss'][k])


@pytest.mark.parametrize('name', cases.LOSS_CASES)
def test_ssim_module_dict_vs_golden(name):
    ML = _mods()
    a, b, f = (T(x).cuda() for x in cases.loss_case(name))
    mod = ML.SSIM(11, 1.0).cuda()
    d1, d2 = mod(a, f), mod(b, f)
    got = np.stack([d1['ssim'].cpu().numpy(), d1['cs'].cpu().numpy(), d1['sigma'].cpu().numpy(),
                    d2['ssim'].cpu().numpy(), d2['cs'].cpu().numpy(), d2['sigma'].cpu().numpy()])
    r32, r64 = LG[f'{name}/f32/ssim_dict'], LG[f'{name}/f64/ssim_dict']
    for i in range(got.shape[0]):
        for j in range(got.shape[1]):
            gates.assert_scalar(f'{name}/dict[{i},{j}]', got[i, j], r32[i, j], r64[i, j])


@pytest.mark.parametrize('name', cases.LOSS_CASES)
def test_loss_backward_vs_fp64_oracle(name):
    """Gradient gate of SURVEY.md 8(c): max-norm error against the fp64 oracle <= 1e-5 max|g|, or —
    where the fp32 reference itself is noisier than that (flat regions: 1/C2 amplifies the rounding
    of the moments) — no worse than the reference's own fp32 error on the same input."""
    a, b, f = (T(x) for x in cases.loss_case(name))
    _, grads = run_new(a, b, f, need_grad=True)
    ref = LG[f'{name}/f64/grad']
    tot64 = ref.sum(axis=0)
    ref32_err = np.abs(LG[f'{name}/f32/grad_total'] - tot64).max() / np.abs(tot64).max()
    frac, mx, where = gates.grad_report(grads[0], ref[0])
    scale_ratio = np.abs(ref[0]).max() / np.abs(tot64).max()
    assert mx <= max(gates.RTOL, ref32_err / max(scale_ratio, 1e-30)), \
        f'{name}: ssim grad max-norm err {mx:.3e} at {where} (fp32 reference: {ref32_err:.3e})'
    # L1 terms: every element is held to 1e-5 max|g| except those a NEAR tie of a sign() argument can flip (masked and
    # counted); an EXACT tie (sign(0) = 0) is asserted, not skipped: with imgf == max(img1, img2) the pixel
    # gradient is exactly zero everywhere.
    a64, b64, f64 = cases.loss_case(name)
    pix_mask, sob_mask, pix_zero = gates.l1_tie_masks(a64, b64, f64)
    for k, nm, mask in ((1, 'pixel', pix_mask), (2, 'grad', sob_mask)):
        frac, mx, masked = gates.masked_grad_report(grads[k], ref[k], mask)
        assert frac <= 1e-4, f'{name}: {nm} grad: {frac:.2e} of the untied elements differ, max {mx:.3e} ({masked:.2e} masked as ties)'
        assert masked <= (1.0 if name in cases.LOSS_GRAD_TIE_CASES else 1e-3), f'{name}: {nm}: {masked:.2e} of the elements are ties'
    assert np.all(grads[1][pix_zero] == 0.0), f'{name}: pixel gradient must be exactly 0 where imgf == max(img1, img2)'
    if name == 'constant':      # constant images: every Sobel response is a structural zero, sign(0) = 0
        assert np.all(grads[2] == 0.0) and np.all(grads[1] != 0.0)
    if name == 'ir_crop_max':   # 8-bit data with imgf = max(img1, img2): 93 % Sobel ties; a flipped sign moves an element by at
        k = 0.1 / f64.size      # most 2 k_grad (|Kx| + |Ky|) = 32 k_grad (64 with the reflect folds)
        assert np.abs(grads[2] - ref[2]).max() <= 64 * k * (1 + 1e-5)


def test_total_backward_and_memo_single_node():
    """train.py:64-71: total = l1+l2+l3; one backward; gradient = sum of the three terms."""
    ML = _mods()
    a, b, f = (T(x) for x in cases.loss_case('rand_3x64x96'))
    A, B_, F_ = a.cuda(), b.cuda(), f.cuda().requires_grad_(True)
    tot = ML.SSIMLoss('ssim', weight=1.0)(A, B_, F_) + ML.PixelLoss('l1', 0.01)(A, B_, F_, mode='max') \
        + ML.GradLoss('l1', 0.1)(A, B_, F_, mode='max')
    tot.backward()
    ref = LG['rand_3x64x96/f64/grad'].sum(axis=0)
    frac, mx, where = gates.grad_report(F_.grad.cpu().numpy(), ref)
    assert frac <= 1e-4, (frac, mx, where)
    gates.assert_scalar('total', tot.item(), LG['rand_3x64x96/f32/loss'].sum(), LG['rand_3x64x96/f64/loss'].sum())


@pytest.mark.parametrize('pixel', [('l1', 'avg'), ('l2', 'max'), ('l2', 'avg')])
def test_secondary_modes(pixel):
    a, b, f = (T(x) for x in cases.loss_case('rand_2x40x37'))
    vals, grads = run_new(a, b, f, True, pixel=pixel, grad=pixel)
    r32 = [OL.pixel_loss(a, b, f, pixel[0], 0.01, pixel[1]).item(), OL.grad_loss(a, b, f, pixel[0], 0.1, pixel[1]).item()]
    ad, bd = a.double(), b.double()
    fd = f.double().requires_grad_(True)
    p64 = OL.pixel_loss(ad, bd, fd, pixel[0], 0.01, pixel[1])
    g64 = OL.grad_loss(ad, bd, fd, pixel[0], 0.1, pixel[1])
    gp, = torch.autograd.grad(p64, fd)
    gg, = torch.autograd.grad(g64, fd)
    gates.assert_scalar('pixel', vals[1], r32[0], p64.item())
    gates.assert_scalar('grad', vals[2], r32[1], g64.item())
    for got, ref, nm in ((grads[1], gp.numpy(), 'pixel'), (grads[2], gg.numpy(), 'grad')):
        frac, mx, where = gates.grad_report(got, ref)
        assert frac <= 1e-4, f'{nm} {pixel}: {frac:.2e} differ, max {mx:.3e} at {where}'


def test_errors_match_reference():
    ML = _mods()
    x = torch.rand(1, 1, 32, 32, device='cuda')
    with pytest.raises(ValueError):
        ML.SSIMLoss('nope')(x, x, x)
    with pytest.raises(ValueError):
        ML.PixelLoss('l3')(x, x, x, mode='max')
    assert ML.PixelLoss('l1')(x, x, x, mode='other') is None
    with pytest.raises(NotImplementedError):
        ML.SSIM(win_size=19)(x, x)                       # windows of 2..17 taps are built
    with pytest.raises(NotImplementedError):
        ML.SSIMLoss('ssim')(x.clone().requires_grad_(True), x, x)   # gradients w.r.t. the sources
    with pytest.raises(Exception):
        ML.SSIMLoss('ssim')(x.cpu(), x.cpu(), x.cpu())   # no CPU fallback


@pytest.mark.parametrize('shape', [(1, 1024, 1224), (2, 517, 1030), (1, 300, 2050)])
def test_loss_config_sizes_vs_live_oracle(shape):
    """BASELINE config 1 shape (1224x1024) and ragged multi-strip / multi-segment shapes."""
    g = torch.Generator().manual_seed(sum(shape))
    a, b, f = (torch.rand((shape[0], 1) + shape[1:], generator=g) for _ in range(3))
    vals, grads = run_new(a, b, f, True)
    r32 = [t.item() for t in OL.train_objective(a, b, f)]
    (l64, g64) = OL.train_objective_grad(a.double(), b.double(), f.double())
    for k, nm in enumerate(('ssim', 'pixel', 'grad')):
        gates.assert_scalar(f'{shape}/{nm}', vals[k], r32[k], l64[k].item())
    tot = grads[0] + grads[1] + grads[2]
    frac, mx, where = gates.grad_report(tot, g64.numpy())
    assert frac <= 1e-4, f'{shape}: total grad {frac:.2e} differ, max {mx:.3e} at {where}'


def test_backward_is_linear_in_upstream_and_deterministic():
    """Size-independent properties at a large size (no oracle needed)."""
    ML = _mods()
    g = torch.Generator().manual_seed(5)
    a, b, f = (torch.rand(2, 1, 1536, 2048, generator=g).cuda() for _ in range(3))
    f.requires_grad_(True)
    l1 = ML.SSIMLoss('ssim')(a, b, f)
    l2 = ML.PixelLoss('l1', 0.01)(a, b, f, mode='max')
    l3 = ML.GradLoss('l1', 0.1)(a, b, f, mode='max')
    parts = [torch.autograd.grad(t, f, retain_graph=True)[0] for t in (l1, l2, l3)]
    tot, = torch.autograd.grad(2.0 * l1 + 3.0 * l2 - 0.5 * l3, f, retain_graph=True)
    ref = 2.0 * parts[0] + 3.0 * parts[1] - 0.5 * parts[2]
    assert (tot - ref).abs().max().item() <= 1e-6 * ref.abs().max().item()
    tot2, = torch.autograd.grad(2.0 * l1 + 3.0 * l2 - 0.5 * l3, f, retain_graph=True)
    assert torch.equal(tot, tot2)
    v1 = ML.SSIMLoss('ssim')(a, b, f.detach().clone()).item()
    assert v1 == l1.item()


def test_single_pass_equals_two_kernel_path():
    """The single-pass forward (loss + gradient in one launch, rescaled in backward) and the
    two-kernel path (forward, recomputing backward) are the same numbers."""
    ML = _mods()
    a, b, f = (T(x) for x in cases.loss_case('rand_2x150x260'))
    res = {}
    for single in (True, False):
        ML.SINGLE_PASS = single
        try:
            A, B_, F_ = a.cuda(), b.cuda(), f.cuda().requires_grad_(True)
            tot = ML.SSIMLoss('ssim', weight=1.0)(A, B_, F_) + ML.PixelLoss('l1', 0.01)(A, B_, F_, mode='max') \
                + ML.GradLoss('l1', 0.1)(A, B_, F_, mode='max')
            (2.5 * tot).backward()
            res[single] = (tot.item(), F_.grad.cpu().numpy())
        finally:
            ML.SINGLE_PASS = True
    assert abs(res[True][0] - res[False][0]) <= 2e-7 * abs(res[False][0])
    scale = np.abs(res[False][1]).max()
    assert np.abs(res[True][1] - res[False][1]).max() <= 1e-6 * scale
    ref = 2.5 * LG['rand_2x150x260/f64/grad'].sum(axis=0)
    frac, mx, where = gates.grad_report(res[True][1], ref)
    assert frac <= 1e-4, (frac, mx, where)


def _grad_vs_oracle(loss_new_fn, loss_ref_fn, a, b, f, rtol=1e-5):
    A, B_, F_ = a.cuda(), b.cuda(), f.cuda().requires_grad_(True)
    ln = loss_new_fn(A, B_, F_)
    gn, = torch.autograd.grad(ln, F_)
    r32 = loss_ref_fn(a, b, f).item()
    fd = f.double().requires_grad_(True)
    l64 = loss_ref_fn(a.double(), b.double(), fd)
    g64, = torch.autograd.grad(l64, fd)
    f32 = f.clone().requires_grad_(True)
    g32, = torch.autograd.grad(loss_ref_fn(a, b, f32), f32)
    gates.assert_scalar('loss', ln.item(), r32, l64.item())
    frac, mx, where = gates.grad_report(gn.cpu().numpy(), g64.numpy(), rtol)
    ref_err = np.abs(g32.numpy() - g64.numpy()).max() / np.abs(g64.numpy()).max()
    assert mx <= max(rtol, ref_err), f'grad max-norm err {mx:.3e} at {where} (fp32 reference: {ref_err:.3e})'


def test_w_ssim_forward_backward():
    ML = _mods()
    a, b, f = (T(x) for x in cases.loss_case('rand_3x64x96'))
    _grad_vs_oracle(lambda x1, x2, y: ML.SSIMLoss('w-ssim', weight=0.7)(x1, x2, y),
                    lambda x1, x2, y: OL.ssim_loss(x1, x2, y, 'w-ssim', weight=0.7), a, b, f)


def test_ms_ssim_forward_backward():
    ML = _mods()
    g = torch.Generator().manual_seed(3)
    a, b, f = (torch.rand(2, 1, 181, 200, generator=g) for _ in range(3))     # odd sizes down the pyramid: 181 -> 91 -> 46 -> 23 -> 12
    f = (0.5 * (a + b) + 0.1 * f).contiguous()
    _grad_vs_oracle(lambda x1, x2, y: ML.SSIMLoss('ms-ssim')(x1, x2, y),
                    lambda x1, x2, y: OL.ssim_loss(x1, x2, y, 'ms-ssim'), a, b, f)
    ms = ML.MS_SSIM()(a.cuda(), f.cuda())
    ref = OL.msssim(a, f, window=OL.window2d(11, 1.5), data_range=1.0)
    np.testing.assert_allclose(ms.cpu().numpy(), ref.numpy(), rtol=1e-5)


def test_use_padding_forward_backward():
    ML = _mods()
    a, b, f = (T(x) for x in cases.loss_case('rand_2x40x37'))
    _grad_vs_oracle(lambda x1, x2, y: ML.SSIMLoss('ssim', use_padding=True)(x1, x2, y),
                    lambda x1, x2, y: OL.ssim_loss(x1, x2, y, 'ssim', use_padding=True), a, b, f)
    d = ML.SSIM(use_padding=True)(a.cuda(), f.cuda())
    ref = OL.ssim(a, f, window=OL.window2d(11, 1.5), data_range=1.0, use_padding=True)
    np.testing.assert_allclose(d['ssim'].cpu().numpy(), ref['ssim'].numpy(), rtol=1e-5)


@pytest.mark.parametrize('mode', ['l1', 'l2'])
def test_tv_loss_forward_backward(mode):
    ML = _mods()
    a, b, f = (T(x) for x in cases.loss_case('rand_2x40x37'))
    x = (f - a).cuda().requires_grad_(True)
    l = ML.TVLoss(mode, weight=0.3)(x)
    g, = torch.autograd.grad(l, x)
    xd = (f - a).double().requires_grad_(True)
    l64 = OL.tv_loss(xd, mode, 0.3)
    g64, = torch.autograd.grad(l64, xd)
    gates.assert_scalar('tv', l.item(), OL.tv_loss(f - a, mode, 0.3).item(), l64.item())
    frac, mx, where = gates.grad_report(g.cpu().numpy(), g64.numpy())
    assert frac <= (1e-3 if mode == 'l1' else 0.0), (frac, mx, where)


def test_msw_ssim_forward_backward():
    ML = _mods()
    for name in ('rand_3x64x96', 'rand_2x40x37'):
        a, b, f = (T(x) for x in cases.loss_case(name))
        _grad_vs_oracle(lambda x1, x2, y: ML.SSIMLoss('msw-ssim', weight=1.3)(x1, x2, y),
                        lambda x1, x2, y: OL.ssim_loss(x1, x2, y, 'msw-ssim', weight=1.3), a, b, f)


def test_ms_ssim_with_padding_pads_every_level():
    ML = _mods()
    g = torch.Generator().manual_seed(8)
    a, b, f = (torch.rand(2, 1, 93, 120, generator=g) for _ in range(3))     # 93 -> 47 -> 24 -> 12 -> 6 (valid windows would fail)
    f = (0.5 * (a + b) + 0.1 * f).contiguous()
    _grad_vs_oracle(lambda x1, x2, y: ML.SSIMLoss('ms-ssim', use_padding=True)(x1, x2, y),
                    lambda x1, x2, y: OL.ssim_loss(x1, x2, y, 'ms-ssim', use_padding=True), a, b, f)
    ms = ML.MS_SSIM(use_padding=True)(a.cuda(), f.cuda())
    ref = OL.msssim(a, f, window=OL.window2d(11, 1.5), data_range=1.0, use_padding=True)
    ref64 = OL.msssim(a.double(), f.double(), window=OL.window2d(11, 1.5).double(), data_range=1.0, use_padding=True)
    for k in range(2):       # a product of five level values: the reference's own fp32 noise reaches 1e-5 here
        gates.assert_scalar(f'ms_ssim_pad[{k}]', ms[k].item(), ref[k].item(), ref64[k].item())


def test_msw_ssim_with_padding():
    ML = _mods()
    a, b, f = (T(x) for x in cases.loss_case('rand_2x40x37'))
    _grad_vs_oracle(lambda x1, x2, y: ML.SSIMLoss('msw-ssim', use_padding=True)(x1, x2, y),
                    lambda x1, x2, y: OL.ssim_loss(x1, x2, y, 'msw-ssim', use_padding=True), a, b, f)


def test_ssim_maps_and_auto_range():
    ML = _mods()
    a, b, f = (T(x) for x in cases.loss_case('rand_2x40x37'))
    d = ML.SSIM(size_average=False, data_range=None)(a.cuda(), f.cuda())
    ref = OL.ssim(a.double(), f.double(), window=OL.window2d(11, 1.5).double(), data_range=None, size_average=False)
    for key in ('ssim', 'cs', 'sigma'):
        assert tuple(d[key].shape) == tuple(ref[key].shape)
        np.testing.assert_allclose(d[key].cpu().numpy(), ref[key].numpy(), rtol=2e-5, atol=2e-6)
    big = (a * 255.0)
    l_new = ML.SSIMLoss('ssim', data_range=None)(big.cuda(), (b * 255.0).cuda(), (f * 255.0).cuda()).item()
    l_ref = OL.ssim_loss(big, b * 255.0, f * 255.0, 'ssim', data_range=None).item()
    assert abs(l_new - l_ref) <= 1e-5 * abs(l_ref) + 1e-7


@pytest.mark.parametrize('mode', ['l1', 'l2'])
def test_norm_loss_forward_backward(mode):
    ML = _mods()
    g = torch.Generator().manual_seed(2)
    x = (torch.rand(3, 2, 33, 65, generator=g) - 0.5)
    xg = x.cuda().requires_grad_(True)
    l = ML.NormLoss(mode, weight=0.4)(xg)
    gx, = torch.autograd.grad(l, xg)
    xd = x.double().requires_grad_(True)
    l64 = OL.norm_loss(xd, mode, 0.4)
    g64, = torch.autograd.grad(l64, xd)
    assert abs(l.item() - l64.item()) <= 1e-6 * abs(l64.item())
    np.testing.assert_allclose(gx.cpu().numpy(), g64.numpy(), rtol=1e-6, atol=1e-12)


@pytest.mark.parametrize('use_padding', [False, True])
def test_ssim_module_dict_is_differentiable(use_padding):
    """SSIM.forward (loss.py:163-185): gradients of the per-sample 'ssim' and 'cs' entries w.r.t. BOTH images,
    against the fp64 oracle's autograd (max-norm gate of the loss gradients)."""
    ML = _mods()
    a, _, f = (T(x) for x in cases.loss_case('rand_3x64x96'))
    w1 = torch.tensor([0.7, -1.3, 2.0])
    w2 = torch.tensor([1.5, 0.25, -0.5])

    def obj(d, dt):
        return (w1.to(d['ssim'].device, dt) * d['ssim']).sum() + (w2.to(d['cs'].device, dt) * d['cs']).sum() + 0.0 * d['sigma'].sum()

    A, F_ = a.cuda().requires_grad_(True), f.cuda().requires_grad_(True)
    d = ML.SSIM(11, 1.0, use_padding).cuda()(A, F_)
    val = (w1.cuda() * d['ssim']).sum() + (w2.cuda() * d['cs']).sum()
    gA, gF = torch.autograd.grad(val, (A, F_))
    a64, f64 = a.double().requires_grad_(True), f.double().requires_grad_(True)
    ref = obj(OL.ssim(a64, f64, 11, None, 1.0, use_padding), torch.float64)
    rA, rF = torch.autograd.grad(ref, (a64, f64))
    a32, f32 = a.clone().requires_grad_(True), f.clone().requires_grad_(True)
    r32 = obj(OL.ssim(a32, f32, 11, None, 1.0, use_padding), torch.float32)
    qA, qF = torch.autograd.grad(r32, (a32, f32))
    gates.assert_scalar('dict objective', val.item(), r32.item(), ref.item())
    for nm, got, r64, q32 in (('d/d img1', gA, rA, qA), ('d/d img2', gF, rF, qF)):
        frac, mx, where = gates.grad_report(got.cpu().numpy(), r64.numpy())
        ref_err = np.abs(q32.numpy() - r64.numpy()).max() / np.abs(r64.numpy()).max()
        assert mx <= max(1e-5, ref_err), f'{nm}: max-norm err {mx:.3e} at {where} (fp32 reference: {ref_err:.3e})'
    # 'sigma' = clamp(var(img1), 1e-4) depends on img1 only: its gradient comes from the generic backward
    w3 = torch.tensor([0.3, 1.1, -0.9])
    d = ML.SSIM(11, 1.0).cuda()(A, F_)
    gA, = torch.autograd.grad((w3.cuda() * d['sigma']).sum() + (w1.cuda() * d['ssim']).sum(), (A,))
    o = OL.ssim(a64, f64, 11, None, 1.0, False)
    rA, = torch.autograd.grad((w3.double() * o['sigma']).sum() + (w1.double() * o['ssim']).sum(), (a64,))
    frac, mx, where = gates.grad_report(gA.cpu().numpy(), rA.numpy())
    assert mx <= 1e-5, f"d(sigma + ssim)/d img1: max-norm err {mx:.3e} at {where}"


@pytest.mark.parametrize('win', [9, 7, 5, 3])
@pytest.mark.parametrize('use_padding', [False, True])
def test_ssim_module_other_window_sizes(win, use_padding):
    """SSIM(win_size=9/7/5/3) (loss.py:163-185; window sigma 0.15 (win-1), loss.py:34): the per-sample dict against the
    oracle, and the gradients of its 'ssim' / 'cs' entries w.r.t. both images against the fp64 oracle's autograd."""
    ML = _mods()
    a, _, f = (T(x) for x in cases.loss_case('rand_3x64x96'))
    w1 = torch.tensor([0.7, -1.3, 2.0])
    w2 = torch.tensor([1.5, 0.25, -0.5])

    def obj(d):
        return (w1.to(d['ssim']) * d['ssim']).sum() + (w2.to(d['cs']) * d['cs']).sum()

    mod = ML.SSIM(win, 1.0, use_padding).cuda()
    assert tuple(mod.window.shape) == (1, 1, win, win)
    with torch.no_grad():
        d0 = mod(a.cuda(), f.cuda())
    o32, o64 = OL.ssim(a, f, win, None, 1.0, use_padding), OL.ssim(a.double(), f.double(), win, None, 1.0, use_padding)
    g32, g64 = WG[f'ssim/win{win}/pad{int(use_padding)}/f32'], WG[f'ssim/win{win}/pad{int(use_padding)}/f64']    # real reference
    for ki, key in enumerate(('ssim', 'cs', 'sigma')):
        for n in range(3):
            gates.assert_scalar(f'win{win}/{key}[{n}]', d0[key][n].item(), o32[key][n].item(), o64[key][n].item())
            gates.assert_scalar(f'win{win}/{key}[{n}] vs golden', d0[key][n].item(), g32[ki, n], g64[ki, n])
    A, F_ = a.cuda().requires_grad_(True), f.cuda().requires_grad_(True)
    val = obj(mod(A, F_))
    gA, gF = torch.autograd.grad(val, (A, F_))
    a64, f64 = a.double().requires_grad_(True), f.double().requires_grad_(True)
    rA, rF = torch.autograd.grad(obj(OL.ssim(a64, f64, win, None, 1.0, use_padding)), (a64, f64))
    a32, f32 = a.clone().requires_grad_(True), f.clone().requires_grad_(True)
    qA, qF = torch.autograd.grad(obj(OL.ssim(a32, f32, win, None, 1.0, use_padding)), (a32, f32))
    for nm, got, r64, q32 in (('d/d img1', gA, rA, qA), ('d/d img2', gF, rF, qF)):
        frac, mx, where = gates.grad_report(got.cpu().numpy(), r64.numpy())
        ref_err = np.abs(q32.numpy() - r64.numpy()).max() / np.abs(r64.numpy()).max()
        # gate: 1e-5 of max|g|, or (3-tap window, sigma 0.3: nearly a delta, tiny variances) 1.5x the band the reference's
        # own fp32 graph keeps around the fp64 gradient — the same rule tests/gates.py applies to scalars
        assert mx <= max(1e-5, 1.5 * ref_err), f'win {win} {nm}: max-norm err {mx:.3e} at {where} (fp32 reference: {ref_err:.3e})'


@pytest.mark.parametrize('win,use_padding', [(7, False), (5, True)])
def test_ms_ssim_module_other_window_sizes(win, use_padding):
    """MS_SSIM(win_size=7/5) (loss.py:188-208 -> calc_msssim, loss.py:113-160, the module's window on every level):
    per-sample values and the gradient w.r.t. the second image against the oracle."""
    ML = _mods()
    g = torch.Generator().manual_seed(77)
    a, f = (torch.rand(2, 1, 208, 240, generator=g) for _ in range(2))
    f = (0.6 * a + 0.4 * f).contiguous()
    w = torch.tensor([1.0, -0.5])
    mod = ML.MS_SSIM(win, 1.0, use_padding).cuda()
    F_ = f.cuda().requires_grad_(True)
    ms = mod(a.cuda(), F_)
    o32 = OL.msssim(a, f, win, None, None, 1.0, use_padding)
    f64 = f.double().requires_grad_(True)
    o64 = OL.msssim(a.double(), f64, win, None, None, 1.0, use_padding)
    assert np.array_equal(a.numpy(), WG['ms/x']) and np.array_equal(f.numpy(), WG['ms/y'])      # the golden inputs
    g32, g64 = WG[f'msssim/win{win}/pad{int(use_padding)}/f32'], WG[f'msssim/win{win}/pad{int(use_padding)}/f64']
    for n in range(2):
        gates.assert_scalar(f'ms-ssim win{win}[{n}]', ms[n].item(), o32[n].item(), o64[n].item())
        gates.assert_scalar(f'ms-ssim win{win}[{n}] vs golden', ms[n].item(), g32[n], g64[n])
    (w.cuda() * ms).sum().backward()
    (w.double() * o64).sum().backward()
    f32 = f.clone().requires_grad_(True)
    (w * OL.msssim(a, f32, win, None, None, 1.0, use_padding)).sum().backward()
    frac, mx, where = gates.grad_report(F_.grad.cpu().numpy(), f64.grad.numpy())
    ref_err = np.abs(f32.grad.numpy() - f64.grad.numpy()).max() / np.abs(f64.grad.numpy()).max()
    assert mx <= max(1e-5, 1.5 * ref_err), f'win {win}: max-norm err {mx:.3e} at {where} (fp32 reference: {ref_err:.3e})'


def test_train_step_shape_through_a_network_matches_the_oracle():
    """train.py:64-71 as a whole: imgf = model(img1, img2) (a non-leaf), the three drop-in losses, backward, clip, Adam.
    The parameter gradients must equal those of the same network under the fp64 oracle loss; also with a NON-contiguous
    imgf (channel 0 of a two-channel output)."""
    ML = _mods()
    torch.manual_seed(3)
    g = torch.Generator().manual_seed(31)
    a, b = (torch.rand(2, 1, 48, 72, generator=g) for _ in range(2))

    def make(dtype):
        torch.manual_seed(3)
        net = torch.nn.Sequential(torch.nn.Conv2d(2, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(4, 2, 3, padding=1))
        return net.to(dtype)

    keep = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # the test network must run in real fp32 to be comparable
    try:
        _train_step_cases(ML, make, a, b)
    finally:
        torch.backends.cudnn.allow_tf32 = keep


def _train_step_cases(ML, make, a, b):
    for slice_channel in (False, True):
        net32 = make(torch.float32).cuda()
        net64 = make(torch.float64)
        A, B_ = a.cuda(), b.cuda()
        out = net32(torch.cat([A, B_], dim=1))
        imgf = out[:, :1] if slice_channel else out.sum(dim=1, keepdim=True)
        assert imgf.is_contiguous() != slice_channel
        loss = ML.SSIMLoss('ssim', weight=1.0)(A, B_, imgf) + ML.PixelLoss('l1', 0.01)(A, B_, imgf, mode='max') \
            + ML.GradLoss('l1', 0.1)(A, B_, imgf, mode='max')
        loss.backward()
        o64 = net64(torch.cat([a, b], dim=1).double())
        f64 = o64[:, :1] if slice_channel else o64.sum(dim=1, keepdim=True)
        l64 = sum(OL.train_objective(a.double(), b.double(), f64))
        l64.backward()
        assert abs(loss.item() - l64.item()) <= 1e-5 * abs(l64.item())
        for p32, p64 in zip(net32.parameters(), net64.parameters()):
            if p64.grad is None:
                assert p32.grad is None or float(p32.grad.abs().max()) == 0.0
                continue
            ref = p64.grad.numpy()
            got = p32.grad.double().cpu().numpy()
            assert np.abs(got - ref).max() <= 2e-4 * max(np.abs(ref).max(), 1e-12), (slice_channel, np.abs(got - ref).max(), np.abs(ref).max())
        opt = torch.optim.Adam(net32.parameters(), lr=1e-4)
        torch.nn.utils.clip_grad_norm_(net32.parameters(), 5.0)
        opt.step()


def _dict_obj(d, ws):
    return sum((w.to(d[k]) * d[k]).sum() for k, w in zip(('ssim', 'cs', 'sigma'), ws))


@pytest.mark.parametrize('win', [13, 8, 4, 2, 17])
@pytest.mark.parametrize('use_padding', [False, True])
def test_ssim_module_any_window_size(win, use_padding):
    """SSIM(win_size) for windows the strip kernels are not instantiated for (odd or even, up to 17 taps; loss.py:33-39
    builds any size): the generic kernels (mmif_ssim_generic_fwd / _bwd).  Dict against the oracle (fp32 band around fp64),
    gradients of ALL THREE entries w.r.t. BOTH images against the fp64 oracle's autograd."""
    ML = _mods()
    a, _, f = (T(x) for x in cases.loss_case('rand_3x64x96'))
    ws = (torch.tensor([0.7, -1.3, 2.0]), torch.tensor([1.5, 0.25, -0.5]), torch.tensor([0.4, 0.9, -1.1]))
    mod = ML.SSIM(win, 1.0, use_padding).cuda()
    assert tuple(mod.window.shape) == (1, 1, win, win)
    with torch.no_grad():
        d0 = mod(a.cuda(), f.cuda())
    o32, o64 = OL.ssim(a, f, win, None, 1.0, use_padding), OL.ssim(a.double(), f.double(), win, None, 1.0, use_padding)
    for key in ('ssim', 'cs', 'sigma'):
        for n in range(3):
            gates.assert_scalar(f'win{win}/{key}[{n}]', d0[key][n].item(), o32[key][n].item(), o64[key][n].item())
    A, F_ = a.cuda().requires_grad_(True), f.cuda().requires_grad_(True)
    gA, gF = torch.autograd.grad(_dict_obj(mod(A, F_), ws), (A, F_))
    a64, f64 = a.double().requires_grad_(True), f.double().requires_grad_(True)
    rA, rF = torch.autograd.grad(_dict_obj(OL.ssim(a64, f64, win, None, 1.0, use_padding), ws), (a64, f64))
    a32, f32 = a.clone().requires_grad_(True), f.clone().requires_grad_(True)
    qA, qF = torch.autograd.grad(_dict_obj(OL.ssim(a32, f32, win, None, 1.0, use_padding), ws), (a32, f32))
    for nm, got, r64, q32 in (('d/d img1', gA, rA, qA), ('d/d img2', gF, rF, qF)):
        frac, mx, where = gates.grad_report(got.cpu().numpy(), r64.numpy())
        ref_err = np.abs(q32.numpy() - r64.numpy()).max() / np.abs(r64.numpy()).max()
        assert mx <= max(1e-5, ref_err), f'win {win} {nm}: max-norm err {mx:.3e} at {where} (fp32 reference: {ref_err:.3e})'
    with pytest.raises(NotImplementedError):
        ML.SSIM(19, 1.0).cuda()(a.cuda(), f.cuda())


@pytest.mark.parametrize('win', [11, 6])
def test_ssim_maps_carry_gradients(win):
    """size_average=False (loss.py:99-108): per-position maps, differentiable w.r.t. both images (the reference's autograd
    gives this for free; here the generic backward with per-position upstream gradients)."""
    ML = _mods()
    a, _, f = (T(x) for x in cases.loss_case('rand_3x64x96'))
    g = torch.Generator().manual_seed(5)
    Ho, Wo = a.shape[-2] - win + 1, a.shape[-1] - win + 1
    wm = [torch.randn(3, 1, Ho, Wo, generator=g) for _ in range(3)]
    mod = ML.SSIM(win, 1.0, False, size_average=False).cuda()
    A, F_ = a.cuda().requires_grad_(True), f.cuda().requires_grad_(True)
    d = mod(A, F_)
    o64 = OL.ssim(a.double(), f.double(), win, None, 1.0, False, size_average=False)
    o32 = OL.ssim(a, f, win, None, 1.0, False, size_average=False)
    for key in ('ssim', 'cs', 'sigma'):
        assert tuple(d[key].shape) == (3, 1, Ho, Wo)
        err = (d[key].detach().cpu().double() - o64[key]).abs().max().item()
        ref = (o32[key].double() - o64[key]).abs().max().item()
        assert err <= max(1e-5 * o64[key].abs().max().item() + 1e-7, ref), f'{key} map: {err:.3e} (fp32 reference {ref:.3e})'
    gA, gF = torch.autograd.grad(_dict_obj(d, wm), (A, F_))
    a64, f64 = a.double().requires_grad_(True), f.double().requires_grad_(True)
    rA, rF = torch.autograd.grad(_dict_obj(OL.ssim(a64, f64, win, None, 1.0, False, size_average=False), wm), (a64, f64))
    a32, f32 = a.clone().requires_grad_(True), f.clone().requires_grad_(True)
    qA, qF = torch.autograd.grad(_dict_obj(OL.ssim(a32, f32, win, None, 1.0, False, size_average=False), wm), (a32, f32))
    for nm, got, r64, q32 in (('d/d img1', gA, rA, qA), ('d/d img2', gF, rF, qF)):
        frac, mx, where = gates.grad_report(got.cpu().numpy(), r64.numpy())
        ref_err = np.abs(q32.numpy() - r64.numpy()).max() / np.abs(r64.numpy()).max()
        assert mx <= max(1e-5, ref_err), f'{nm}: max-norm err {mx:.3e} at {where} (fp32 reference: {ref_err:.3e})'


def test_calc_ssim_function_form_shrinks_the_window():
    """calc_ssim / calc_msssim called as functions (loss.py:52-160): without a window the reference builds one of
    min(win_size, h, w) taps — a 7 x 40 image gets a 7-tap window (sigma 0.9) — and a passed create_window() is honoured."""
    ML = _mods()
    g = torch.Generator().manual_seed(9)
    a, f = torch.rand(2, 1, 7, 40, generator=g), torch.rand(2, 1, 7, 40, generator=g)
    d = ML.calc_ssim(a.cuda(), f.cuda(), data_range=1.0)
    o32, o64 = OL.ssim(a, f, 11, None, 1.0), OL.ssim(a.double(), f.double(), 11, None, 1.0)
    for key in ('ssim', 'cs', 'sigma'):
        for n in range(2):
            gates.assert_scalar(f'shrunk/{key}[{n}]', d[key][n].item(), o32[key][n].item(), o64[key][n].item())
    a, f = torch.rand(2, 1, 6, 40, generator=g), torch.rand(2, 1, 6, 40, generator=g)      # even shrunk window: 6 taps
    d = ML.calc_ssim(a.cuda(), f.cuda(), data_range=1.0)
    o32, o64 = OL.ssim(a, f, 11, None, 1.0), OL.ssim(a.double(), f.double(), 11, None, 1.0)
    for n in range(2):
        gates.assert_scalar(f'shrunk6/ssim[{n}]', d['ssim'][n].item(), o32['ssim'][n].item(), o64['ssim'][n].item())
    x, y = (T(v) for v in cases.loss_case('rand_3x64x96')[::2])
    d = ML.calc_ssim(x.cuda(), y.cuda(), window=ML.create_window(9).cuda(), data_range=1.0)
    o32, o64 = OL.ssim(x, y, 9, None, 1.0), OL.ssim(x.double(), y.double(), 9, None, 1.0)
    for n in range(3):
        gates.assert_scalar(f'window9/ssim[{n}]', d['ssim'][n].item(), o32['ssim'][n].item(), o64['ssim'][n].item())
    with pytest.raises(NotImplementedError):
        ML.calc_ssim(x.cuda(), y.cuda(), window=torch.ones(1, 1, 5, 5) / 25.0, data_range=1.0)
    with pytest.raises(RuntimeError):
        ML.SSIM(11, 1.0).cuda()(a.cuda(), f.cuda())       # the module keeps its 11-tap window: conv2d fails in the reference too
    g = torch.Generator().manual_seed(77)
    p, q = (torch.rand(2, 1, 208, 240, generator=g) for _ in range(2))
    q = (0.6 * p + 0.4 * q).contiguous()
    ms = ML.calc_msssim(p.cuda(), q.cuda(), data_range=1.0)
    o32, o64 = OL.msssim(p, q, 11, None, None, 1.0), OL.msssim(p.double(), q.double(), 11, None, None, 1.0)
    for n in range(2):
        gates.assert_scalar(f'calc_msssim[{n}]', ms[n].item(), o32[n].item(), o64[n].item())


def test_ms_ssim_module_generic_window():
    """MS_SSIM(win_size=13) (loss.py:188-208): the module's 13-tap window on every pyramid level through the generic kernels,
    value and gradient w.r.t. the second image."""
    ML = _mods()
    g = torch.Generator().manual_seed(78)
    a, f = (torch.rand(2, 1, 224, 240, generator=g) for _ in range(2))
    f = (0.6 * a + 0.4 * f).contiguous()
    w = torch.tensor([1.0, -0.5])
    mod = ML.MS_SSIM(13, 1.0).cuda()
    F_ = f.cuda().requires_grad_(True)
    ms = mod(a.cuda(), F_)
    o32 = OL.msssim(a, f, 13, None, None, 1.0)
    f64 = f.double().requires_grad_(True)
    o64 = OL.msssim(a.double(), f64, 13, None, None, 1.0)
    for n in range(2):
        gates.assert_scalar(f'ms-ssim win13[{n}]', ms[n].item(), o32[n].item(), o64[n].item())
    (w.cuda() * ms).sum().backward()
    (w.double() * o64).sum().backward()
    f32 = f.clone().requires_grad_(True)
    (w * OL.msssim(a, f32, 13, None, None, 1.0)).sum().backward()
    frac, mx, where = gates.grad_report(F_.grad.cpu().numpy(), f64.grad.numpy())
    ref_err = np.abs(f32.grad.numpy() - f64.grad.numpy()).max() / np.abs(f64.grad.numpy()).max()
    assert mx <= max(1e-5, ref_err), f'max-norm err {mx:.3e} at {where} (fp32 reference: {ref_err:.3e})'


@pytest.mark.parametrize('use_padding', [False, True])
def test_msw_ssim_size_average_true(use_padding):
    """MSW_SSIM(size_average=True) (loss.py:211-237 with per-sample dict entries): gamma_b from the window-mean clamped
    variances of the sources, one weighted SSIM per window size; also with a window set the strip kernels do not cover."""
    ML = _mods()
    a, b, f = (T(x) for x in cases.loss_case('rand_3x64x96'))

    def ref(wins):
        def fn(x1, x2, y):
            acc = 0.0
            for k in wins:
                o1 = OL.ssim(x1, y, k, None, 1.0, use_padding, size_average=True)
                o2 = OL.ssim(x2, y, k, None, 1.0, use_padding, size_average=True)
                gm = o1['sigma'] / (o1['sigma'] + o2['sigma']).clamp(min=1e-7)
                acc = acc + (gm * o1['ssim']).mean() + ((1.0 - gm) * o2['ssim']).mean()
            return acc / len(wins)
        return fn

    for wins in ((11, 9, 7, 5, 3), (13, 6)):
        _grad_vs_oracle(lambda x1, x2, y: ML.MSW_SSIM(wins, 1.0, use_padding, size_average=True)(x1, x2, y), ref(wins), a, b, f)
