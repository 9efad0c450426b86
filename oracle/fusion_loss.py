"""Oracle restatement of the reference fusion objective (``core/loss.py``).

Functional restatement with torch CPU ops.  Every function cites the reference
lines it follows.  Dtype-generic: windows/filters follow the input dtype, so
``x.double()`` inputs give the float64 oracle.

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.
"""
import math

import torch
import torch.nn.functional as F

MS_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)  # loss.py:122,198
GAMMA_EPS = 1e-7  # loss.py:21


def gauss_taps(win_size, sigma):
    """1-D taps: python-double exp -> float32 tensor -> float32 normalise (loss.py:24-30)."""
    c = win_size // 2
    taps = torch.tensor([math.exp(-(i - c) ** 2 / (2.0 * sigma ** 2)) for i in range(win_size)],
                        dtype=torch.float32)
    return taps / taps.sum()


def loss_sigma(win_size):
    """Window-size -> sigma rule of the loss module (loss.py:34)."""
    return 1.5 if win_size == 11 else 0.15 * (win_size - 1)


def window2d(win_size, sigma):
    """(1,1,k,k) float32 outer-product window (loss.py:36-39, metric.py:299-303)."""
    col = gauss_taps(win_size, sigma).unsqueeze(1)
    return torch.mm(col, col.t())[None, None]


def blur(img, window, use_padding=False):
    """Depthwise *valid* correlation, optional reflect pad k//2 (loss.py:42-49)."""
    if use_padding:
        p = window.shape[-1] // 2
        img = F.pad(img, (p, p, p, p), 'reflect')
    return F.conv2d(img, window, groups=img.shape[1])


def ssim_maps(x, y, window, data_range, use_padding=False):
    """SSIM / CS / clamped-variance maps (loss.py:73-103)."""
    window = window.to(x)
    a, b = x.clone(), y.clone()
    mu_a, mu_b = blur(a, window, use_padding), blur(b, window, use_padding)
    aa, bb, ab = mu_a * mu_a, mu_b * mu_b, mu_a * mu_b
    var_a = (blur(a * a, window, use_padding) - aa).clamp(min=0)
    var_b = (blur(b * b, window, use_padding) - bb).clamp(min=0)
    cov = blur(a * b, window, use_padding) - ab
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    lum_n, lum_d = 2.0 * ab + c1, aa + bb + c1
    str_n, str_d = 2.0 * cov + c2, var_a + var_b + c2
    return {'ssim': (lum_n * str_n) / (lum_d * str_d), 'cs': str_n / str_d,
            'sigma': var_a.clamp(min=1e-4)}


def detect_range(x):
    """data_range=None auto-detect (loss.py:60-63)."""
    hi = 255.0 if x.max() > 128 else 1.0
    lo = -1.0 if x.min() < -0.5 else 0.0
    return hi - lo


def ssim(x, y, win_size=11, window=None, data_range=None, use_padding=False, size_average=True):
    """loss.py:52-110 — returns dict of per-sample means (or maps)."""
    L = detect_range(x) if data_range is None else data_range
    if window is None:
        k = min(win_size, x.shape[-2], x.shape[-1])
        window = window2d(k, loss_sigma(k))
    out = ssim_maps(x, y, window, L, use_padding)
    if size_average:
        out = {k: v.mean(dim=(1, 2, 3)) for k, v in out.items()}
    return out


def halve(img):
    """Odd-size reflect pad (right/bottom) then 2x2 mean pool (loss.py:147-153)."""
    h, w = img.shape[-2:]
    img = F.pad(img, (0, w % 2, 0, h % 2), 'reflect')
    return F.avg_pool2d(img, 2, 2)


def msssim(x, y, win_size=11, window=None, weights=None, data_range=None, use_padding=False,
           size_average=True):
    """loss.py:113-160 — five dyadic levels; cs on 0..3, ssim on 4; prod(v**w)."""
    wts = torch.tensor(MS_WEIGHTS, dtype=torch.float32) if weights is None else weights
    if window is None:
        k = min(win_size, x.shape[-2], x.shape[-1])
        window = window2d(k, loss_sigma(k))
    wts = wts.to(x)
    a, b = x.clone(), y.clone()
    vals = []
    n = len(wts)
    for lvl in range(n):
        o = ssim(a, b, win_size, window, data_range, use_padding, size_average)
        if lvl < n - 1:
            vals.append(o['cs'])
            a, b = halve(a), halve(b)
        else:
            vals.append(o['ssim'])
    vals = torch.stack(vals, dim=0).clamp(min=GAMMA_EPS)
    return torch.prod(vals ** wts.unsqueeze(1), dim=0)


def _gamma(o1, o2):
    return o1['sigma'] / (o1['sigma'] + o2['sigma']).clamp_(min=GAMMA_EPS)


def mswssim(x1, x2, f, win_sizes=(11, 9, 7, 5, 3), data_range=1.0, use_padding=False):
    """loss.py:211-237 — per-pixel variance-weighted SSIM over five window sizes."""
    acc = 0.0
    for k in win_sizes:
        w = window2d(k, loss_sigma(k))
        o1 = ssim(x1, f, window=w, data_range=data_range, use_padding=use_padding, size_average=False)
        o2 = ssim(x2, f, window=w, data_range=data_range, use_padding=use_padding, size_average=False)
        g = _gamma(o1, o2)
        acc = acc + (g * o1['ssim']).mean() + ((1.0 - g) * o2['ssim']).mean()
    return acc / len(win_sizes)


def ssim_loss(x1, x2, f, mode='ssim', data_range=1.0, use_padding=False, weight=1.0):
    """SSIMLoss.forward (loss.py:252-284)."""
    if mode == 'ssim':
        w = window2d(11, loss_sigma(11))
        s1 = ssim(x1, f, window=w, data_range=data_range, use_padding=use_padding)['ssim'].mean()
        s2 = ssim(x2, f, window=w, data_range=data_range, use_padding=use_padding)['ssim'].mean()
        val = (s1 + s2) * 0.5
    elif mode == 'w-ssim':
        w = window2d(11, loss_sigma(11))
        o1 = ssim(x1, f, window=w, data_range=data_range, use_padding=use_padding)
        o2 = ssim(x2, f, window=w, data_range=data_range, use_padding=use_padding)
        g = _gamma(o1, o2)
        val = (g * o1['ssim']).mean() + ((1.0 - g) * o2['ssim']).mean()
    elif mode == 'ms-ssim':
        w = window2d(11, loss_sigma(11))
        m1 = msssim(x1, f, window=w, data_range=data_range, use_padding=use_padding).mean()
        m2 = msssim(x2, f, window=w, data_range=data_range, use_padding=use_padding).mean()
        val = (m1 + m2) * 0.5
    elif mode == 'msw-ssim':
        val = mswssim(x1, x2, f, data_range=data_range, use_padding=use_padding)
    else:
        raise ValueError("only supported ['ssim', 'w-ssim', 'ms-ssim', 'msw-ssim'] mode")
    return weight * (1.0 - val)


def norm_loss(x, mode='l1', weight=1.0):
    """NormLoss.forward (loss.py:375-385)."""
    if mode == 'l1':
        v = torch.abs(x).mean()
    elif mode == 'l2':
        v = torch.pow(x, 2).mean()
    else:
        raise ValueError("only supported ['l1', 'l2'] mode")
    return weight * v


def _combine(t1, t2, tf, norm_mode, weight, mode):
    """avg / max source combination shared by PixelLoss and GradLoss (loss.py:294-304,335-344)."""
    if mode == 'avg':
        return (norm_loss(tf - t1, norm_mode, weight) + norm_loss(tf - t2, norm_mode, weight)) * 0.5
    if mode == 'max':
        return norm_loss(tf - torch.max(t1, t2), norm_mode, weight)
    return None  # the reference silently returns None for any other mode (loss.py:294-304)


def pixel_loss(x1, x2, f, norm_mode='l1', weight=1.0, mode='avg'):
    """PixelLoss.forward (loss.py:294-304)."""
    return _combine(x1, x2, f, norm_mode, weight, mode)


SOBEL_X = ((-1., 0., 1.), (-2., 0., 2.), (-1., 0., 1.))  # loss.py:314-315
SOBEL_Y = ((-1., -2., -1.), (0., 0., 0.), (1., 2., 1.))  # loss.py:316-317


def sobel_l1(img):
    """|Kx * pad_reflect(u)| + |Ky * pad_reflect(u)| (loss.py:322-328)."""
    kx = torch.tensor(SOBEL_X, dtype=torch.float32).reshape(1, 1, 3, 3)
    ky = torch.tensor(SOBEL_Y, dtype=torch.float32).reshape(1, 1, 3, 3)
    p = F.pad(img.clone(), (1, 1, 1, 1), 'reflect')
    return torch.abs(F.conv2d(p, kx.to(p))) + torch.abs(F.conv2d(p, ky.to(p)))


def grad_loss(x1, x2, f, norm_mode='l1', weight=1.0, mode='avg'):
    """GradLoss.forward (loss.py:330-344)."""
    return _combine(sobel_l1(x1), sobel_l1(x2), sobel_l1(f), norm_mode, weight, mode)


def tv_loss(x, norm_mode='l1', weight=1.0):
    """TVLoss.forward (loss.py:354-358)."""
    dv = x[..., 1:, :] - x[..., :-1, :]
    dh = x[..., :, 1:] - x[..., :, :-1]
    return norm_loss(dv, norm_mode, weight) + norm_loss(dh, norm_mode, weight)


def train_objective(x1, x2, f, w_ssim=1.0, w_pixel=0.01, w_grad=0.1):
    """The three terms exactly as train.py:64-69,302-317 wires them."""
    l1 = ssim_loss(x1, x2, f, 'ssim', weight=w_ssim)
    l2 = pixel_loss(x1, x2, f, 'l1', w_pixel, mode='max')
    l3 = grad_loss(x1, x2, f, 'l1', w_grad, mode='max')
    return l1, l2, l3


def train_objective_grad(x1, x2, f, upstream=(1.0, 1.0, 1.0), **kw):
    """Losses and dL/d(imgf) via torch autograd on the restated graph (train.py:69-71)."""
    f = f.detach().clone().requires_grad_(True)
    l1, l2, l3 = train_objective(x1, x2, f, **kw)
    total = upstream[0] * l1 + upstream[1] * l2 + upstream[2] * l3
    total.backward()
    return (l1.detach(), l2.detach(), l3.detach()), f.grad
