"""Validates the kernel-side formulas (tests/kernel_math.py) against the oracle in float64."""
import numpy as np
import pytest
import torch

import cases
import kernel_math as KM
from oracle import fusion_loss as OL

LG = np.load(cases.HERE + '/loss_golden.npz')


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x)).double()


@pytest.mark.parametrize('name', ['rand_2x40x37', 'rand_1x11x11', 'natural_2x96x128', 'unbounded_f', 'polar_crop_avg'])
def test_shifted_forward_matches_reference(name):
    a, b, f = (T(x) for x in cases.loss_case(name))
    s, cs, sg = KM.ssim_forward(a, f)
    ref = OL.ssim(a, f, window=OL.window2d(11, 1.5), data_range=1.0)
    np.testing.assert_allclose(s.mean(dim=(1, 2, 3)).numpy(), ref['ssim'].numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(cs.mean(dim=(1, 2, 3)).numpy(), ref['cs'].numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(sg.mean(dim=(1, 2, 3)).numpy(), ref['sigma'].numpy(), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('name', ['rand_2x40x37', 'rand_1x11x11', 'rand_1x12x300', 'natural_2x96x128', 'unbounded_f'])
def test_gradients_match_autograd(name):
    a, b, f = (T(x) for x in cases.loss_case(name))
    ref = LG[f'{name}/f64/grad']
    g = KM.ssim_loss_grad(a, b, f)
    scale = np.abs(ref[0]).max()
    assert np.abs(g.numpy() - ref[0]).max() <= 2e-6 * scale   # separable taps vs the fp32-rounded 2-D window
    np.testing.assert_allclose(KM.pixel_loss_grad(a, b, f).numpy(), ref[1], rtol=1e-12, atol=1e-18)
    np.testing.assert_allclose(KM.grad_loss_grad(a, b, f).numpy(), ref[2], rtol=1e-9, atol=1e-15)


@pytest.mark.parametrize('combine', ['max', 'avg'])
@pytest.mark.parametrize('norm', [1, 2])
def test_secondary_mode_gradients(combine, norm):
    a, b, f = (T(x) for x in cases.loss_case('rand_2x40x37'))
    fr = f.clone().requires_grad_(True)
    nm = 'l1' if norm == 1 else 'l2'
    OL.grad_loss(a, b, fr, nm, 0.1, combine).backward()
    np.testing.assert_allclose(KM.grad_loss_grad(a, b, f, 0.1, combine, norm).numpy(), fr.grad.numpy(),
                               rtol=1e-9, atol=1e-15)
    fr = f.clone().requires_grad_(True)
    OL.pixel_loss(a, b, fr, nm, 0.01, combine).backward()
    np.testing.assert_allclose(KM.pixel_loss_grad(a, b, f, 0.01, combine, norm).numpy(), fr.grad.numpy(),
                               rtol=1e-12, atol=1e-18)
