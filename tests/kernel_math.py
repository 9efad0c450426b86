"""Torch (CPU) statement of the arithmetic the CUDA kernels use, written the way the kernels
compute it (separable taps, mean-shifted moments with the window-sum correction, adjoint blur of
coefficient maps, folded Sobel adjoint).  ``test_kernel_math.py`` checks it against the oracle in
float64, so the formulas are validated without a GPU; the CUDA code transcribes these."""
import math

import torch
import torch.nn.functional as F

from oracle import fusion_loss as OL


def window_consts(win=11, sigma=1.5):
    taps = OL.gauss_taps(win, sigma)                     # float32 1-D taps
    w2 = torch.mm(taps[:, None], taps[None, :])          # float32 2-D window the reference uses
    s = w2.double().sum().item()                         # sum of the reference window
    return taps, s, s - 1.0


def sep_blur(z, taps):
    k = taps.numel()
    t = taps.to(z)
    return F.conv2d(F.conv2d(z, t.view(1, 1, 1, k)), t.view(1, 1, k, 1))


def shifted_moments(x, y, taps, s, eps, cx, cy):
    """Moments of (x-cx, y-cy) and the reference-equivalent variances/covariance."""
    xs, ys = x - cx, y - cy
    mx, my = sep_blur(xs, taps), sep_blur(ys, taps)
    vx = sep_blur(xs * xs, taps) - mx * mx - eps * cx * (2 * mx + cx * s)
    vy = sep_blur(ys * ys, taps) - my * my - eps * cy * (2 * my + cy * s)
    cov = sep_blur(xs * ys, taps) - mx * my - eps * (cx * my + cy * mx + cx * cy * s)
    return xs, ys, mx, my, vx, vy, cov


def ssim_forward(x, y, L=1.0, win=11, sigma=1.5, cx=None, cy=None):
    taps, s, eps = window_consts(win, sigma)
    cx = x[..., x.shape[-2] // 2, x.shape[-1] // 2].reshape(-1, 1, 1, 1) if cx is None else cx
    cy = y[..., y.shape[-2] // 2, y.shape[-1] // 2].reshape(-1, 1, 1, 1) if cy is None else cy
    xs, ys, mx, my, vx, vy, cov = shifted_moments(x, y, taps, s, eps, cx, cy)
    mux, muy = mx + s * cx, my + s * cy
    vxc, vyc = vx.clamp(min=0), vy.clamp(min=0)
    C1, C2 = (0.01 * L) ** 2, (0.03 * L) ** 2
    A1, B1 = 2 * mux * muy + C1, mux * mux + muy * muy + C1
    A2, B2 = 2 * cov + C2, vxc + vyc + C2
    return (A1 * A2) / (B1 * B2), A2 / B2, vxc.clamp(min=1e-4)


def ssim_coef(x, y, L, taps, s, eps, cx, cy):
    """Per-window partial derivatives w.r.t. the shifted moments of y (DESIGN.md 'backward')."""
    xs, ys, mx, my, vx, vy, cov = shifted_moments(x, y, taps, s, eps, cx, cy)
    mux, muy = mx + s * cx, my + s * cy
    my_mask = (vy >= 0).to(x.dtype)
    vxc, vyc = vx.clamp(min=0), vy.clamp(min=0)
    C1, C2 = (0.01 * L) ** 2, (0.03 * L) ** 2
    A1, B1 = 2 * mux * muy + C1, mux * mux + muy * muy + C1
    A2, B2 = 2 * cov + C2, vxc + vyc + C2
    S = (A1 * A2) / (B1 * B2)
    dcov = 2 * A1 / (B1 * B2)                      # dS/d cov
    dvar = -my_mask * S / B2                       # dS/d var_y (through the clamp)
    dmu = 2 * mux * A2 / (B1 * B2) - 2 * muy * S / B1       # luminance path
    a = dmu - dvar * (2 * my + 2 * eps * cy) - dcov * (mx + eps * cx)
    return a, dvar, dcov, xs, ys


def adjoint_blur(z, taps):
    """Adjoint of the valid separable correlation = full correlation with the flipped taps."""
    k = taps.numel()
    t = taps.to(z).flip(0)
    z = F.pad(z, (k - 1, k - 1, k - 1, k - 1))
    return F.conv2d(F.conv2d(z, t.view(1, 1, 1, k)), t.view(1, 1, k, 1))


def ssim_loss_grad(x1, x2, y, L=1.0, w_ssim=1.0, win=11, sigma=1.5):
    """d/dy of w*(1 - (mean_b ssim(x1,y) + mean_b ssim(x2,y))/2), reference loss.py:253-257."""
    taps, s, eps = window_consts(win, sigma)
    ctr = lambda z: z[..., z.shape[-2] // 2, z.shape[-1] // 2].reshape(-1, 1, 1, 1)
    c1, c2, cy = ctr(x1), ctr(x2), ctr(y)
    a1, b1, g1, x1s, ys = ssim_coef(x1, y, L, taps, s, eps, c1, cy)
    a2, b2, g2, x2s, _ = ssim_coef(x2, y, L, taps, s, eps, c2, cy)
    tot = adjoint_blur(a1 + a2, taps) + 2 * ys * adjoint_blur(b1 + b2, taps) \
        + x1s * adjoint_blur(g1, taps) + x2s * adjoint_blur(g2, taps)
    B, _, H, W = y.shape
    n = B * (H - win + 1) * (W - win + 1)
    return -w_ssim * 0.5 / n * tot


def sobel_xy(u):
    """gx, gy with reflect borders, separable form used by the kernels."""
    p = F.pad(u, (1, 1, 1, 1), 'reflect')
    d = p[..., :, 2:] - p[..., :, :-2]                       # horizontal difference, rows padded
    sm = p[..., :, :-2] + 2 * p[..., :, 1:-1] + p[..., :, 2:]
    gx = d[..., :-2, :] + 2 * d[..., 1:-1, :] + d[..., 2:, :]
    gy = sm[..., 2:, :] - sm[..., :-2, :]
    return gx, gy


def norm_deriv(D, norm):
    return torch.sign(D) if norm == 1 else 2 * D


def grad_loss_grad(x1, x2, y, weight=0.1, combine='max', norm=1):
    """d/dy of GradLoss (loss.py:330-344) by the folded adjoint the bwd kernel uses."""
    B, _, H, W = y.shape
    gx, gy = sobel_xy(y)
    Sy = gx.abs() + gy.abs()
    S1 = sum(v.abs() for v in sobel_xy(x1))
    S2 = sum(v.abs() for v in sobel_xy(x2))
    n = B * H * W
    if combine == 'max':
        r = weight / n * norm_deriv(Sy - torch.max(S1, S2), norm)
    else:
        r = 0.5 * weight / n * (norm_deriv(Sy - S1, norm) + norm_deriv(Sy - S2, norm))
    tx, ty = r * torch.sign(gx), r * torch.sign(gy)
    # G on the padded grid [-1..H] x [-1..W]: G(p') = sum_{a,b} Kx[a,b] tx(p'-(a-1,b-1)) + Ky[a,b] ty(...)
    tx2 = F.pad(tx, (2, 2, 2, 2))
    ty2 = F.pad(ty, (2, 2, 2, 2))   # index (i+2, j+2) <-> q=(i,j); padded grid p' in [-1..H] -> offset +2
    Hp, Wp = H + 2, W + 2
    G = torch.zeros(B, 1, Hp, Wp, dtype=y.dtype)
    KX = OL.SOBEL_X
    KY = OL.SOBEL_Y
    for a in range(3):
        for b in range(3):
            # q = p' - (a-1, b-1); p' index pi in [0,Hp) <-> p' = pi-1; q index in tx2 = (pi-1)-(a-1)+2 = pi - a + 2
            sl = (Ellipsis, slice(2 - a, 2 - a + Hp), slice(2 - b, 2 - b + Wp))
            G = G + KX[a][b] * tx2[sl] + KY[a][b] * ty2[sl]
    out = G[..., 1:-1, 1:-1].clone()
    out[..., 1, :] += G[..., 0, 1:-1]            # row -1 folds onto row 1
    out[..., H - 2, :] += G[..., Hp - 1, 1:-1]   # row H folds onto row H-2
    out[..., :, 1] += G[..., 1:-1, 0]
    out[..., :, W - 2] += G[..., 1:-1, Wp - 1]
    out[..., 1, 1] += G[..., 0, 0]
    out[..., 1, W - 2] += G[..., 0, Wp - 1]
    out[..., H - 2, 1] += G[..., Hp - 1, 0]
    out[..., H - 2, W - 2] += G[..., Hp - 1, Wp - 1]
    return out


def pixel_loss_grad(x1, x2, y, weight=0.01, combine='max', norm=1):
    n = y.numel()
    if combine == 'max':
        return weight / n * norm_deriv(y - torch.max(x1, x2), norm)
    return 0.5 * weight / n * (norm_deriv(y - x1, norm) + norm_deriv(y - x2, norm))
