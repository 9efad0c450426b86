"""Oracle restatement of the reference objective metric suite (``core/metric.py``)
and of the per-pair composition / aggregation in ``eval.py``.

Plain torch (+ numpy for the joint histogram, as the reference) on CPU.  Each
function cites the reference lines it follows.  Dtype-generic except where the
reference itself fixes a dtype (float64 joint histogram, metric.py:141-145).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .fusion_loss import gauss_taps, blur, halve, MS_WEIGHTS, SOBEL_X, SOBEL_Y

METRIC_NAMES = ('sd', 'ag', 'sf', 'mse', 'psnr', 'cc', 'scd', 'en', 'ce', 'mi',
                'qabf', 'nabf', 'labf', 'ssim', 'msssim', 'viff')  # eval.py:52-68


# ---- first/second-order statistics ------------------------------------------------
def mean(img):
    """metric.py:25-26."""
    return img.mean()


def std(img):
    """Population std, two-pass (metric.py:30-34)."""
    d = img.clone()
    d -= d.mean()
    return d.pow(2).mean().pow(0.5)


def avg_gradient(img):
    """metric.py:38-46 — forward differences on the (H-1)x(W-1) grid."""
    u = img.clone()
    base = u[..., :-1, :-1]
    dx = u[..., :-1, 1:] - base
    dy = u[..., 1:, :-1] - base
    return ((dx.pow(2) + dy.pow(2)) * 0.5).pow(0.5).mean()


def spatial_freq(img):
    """metric.py:50-59."""
    u = img.clone()
    dv = u[..., 1:, :] - u[..., :-1, :]
    dh = u[..., :, 1:] - u[..., :, :-1]
    return (dv.pow(2).mean() + dh.pow(2).mean()).pow(0.5)


def mse(a, b):
    """metric.py:63-68 — both images divided by 255 first."""
    e = a.clone() / 255.0 - b.clone() / 255.0
    return e.pow(2).mean()


def psnr(mse_val, L=1.0, root=False):
    """metric.py:72-76."""
    if root:
        return 20.0 * torch.log10(L / mse_val ** 0.5)
    return 10.0 * torch.log10(L ** 2 / mse_val)


def corrcoef(a, b):
    """Pearson r with two-pass centring (metric.py:80-91)."""
    u, v = a.clone(), b.clone()
    u -= u.mean()
    v -= v.mean()
    return (u * v).sum() / ((u * u).sum() * (v * v).sum()).pow(0.5)


def scd(a, b, f):
    """metric.py:95-99."""
    return corrcoef(f - a, b) + corrcoef(f - b, a)


# ---- histogram family ---------------------------------------------------------------
def hist_counts(img):
    """256-bin counts over [0,256] as torch.histc returns them (metric.py:112-113)."""
    return torch.histc(img.clone(), 256, 0, 256)


def joint_counts(a, b):
    """256x256 float64 counts via np.histogram2d (metric.py:139-143)."""
    h = np.histogram2d(a.clone().numpy().flatten(), b.clone().numpy().flatten(), 256,
                       ((0, 256), (0, 256)))[0]
    return torch.from_numpy(h)


def prob(img):
    """metric.py:103-116."""
    return hist_counts(img) / img.numel()


def joint_prob(a, b):
    """metric.py:129-145."""
    return joint_counts(a, b) / a.numel()


def _plogp_sum(p):
    nz = torch.where(p != 0)
    return (-p[nz] * torch.log2(p[nz])).sum()


def entropy(img):
    """metric.py:119-125."""
    return _plogp_sum(prob(img))


def joint_entropy(a, b):
    """metric.py:148-154 (float64)."""
    return _plogp_sum(joint_prob(a, b))


def cross_entropy(a, b):
    """metric.py:158-165 — sum p1 log2(p1/p2) over bins where p1*p2 != 0."""
    p, q = prob(a), prob(b)
    nz = torch.where(p * q != 0)
    return (p[nz] * torch.log2(p[nz] / q[nz])).sum()


def mutual_info(a, b, normalized=False):
    """metric.py:169-188 — en1+en2-en12, result float64."""
    e1, e2, e12 = entropy(a), entropy(b), joint_entropy(a, b)
    mi = e1 + e2 - e12
    return 2.0 * mi / (e1 + e2) if normalized else mi


# ---- edge-preservation family -------------------------------------------------------
def sobel_polar(img):
    """Edge strength sqrt(gx^2+gy^2) and orientation atan2(gy,gx) (metric.py:192-206)."""
    kx = torch.tensor(SOBEL_X, dtype=torch.float32).reshape(1, 1, 3, 3)
    ky = torch.tensor(SOBEL_Y, dtype=torch.float32).reshape(1, 1, 3, 3)
    p = F.pad(img.clone(), (1, 1, 1, 1), 'reflect')
    gx, gy = F.conv2d(p, kx.to(p)), F.conv2d(p, ky.to(p))
    return (gx.pow(2) + gy.pow(2)).pow(0.5), torch.atan2(gy, gx)


QABF_CONST = {'qabf': ((0.9994, 15, 0.5), (0.9879, 22, 0.8)),   # metric.py:217-219
              'nabf': ((0.9999, 19, 0.5), (0.9995, 22, 0.5))}   # metric.py:220-222


def edge_preservation(src, fused, mode='qabf'):
    """calc_Qxy (metric.py:209-230): returns Q = Qg*Qa, g_src, g_fused."""
    gs, as_ = sobel_polar(src)
    gf, af = sobel_polar(fused)
    G = torch.min(gs, gf) / torch.max(gs, gf)
    G[G != G] = 0.0
    A = torch.abs(torch.abs(as_ - af) - math.pi / 2) * 2 / math.pi
    (Gg, kg, sg), (Ga, ka, sa) = QABF_CONST[mode]
    Qg = Gg / (1 + torch.exp(-kg * (G - sg)))
    Qa = Ga / (1 + torch.exp(-ka * (A - sa)))
    return Qg * Qa, gs, gf


def qabf(a, b, f, L=1.5, full=False):
    """calc_Qabf (metric.py:233-256)."""
    Qa, ga, gf = edge_preservation(a, f)
    Qb, gb, _ = edge_preservation(b, f)
    wa, wb = ga.pow(L), gb.pow(L)
    den = (wa + wb).sum()
    q = (Qa * wa + Qb * wb).sum() / den
    if not full:
        return q
    loss_term = (1.0 - Qa) * wa + (1.0 - Qb) * wb
    art = torch.where(gf > torch.max(ga, gb), 1.0, 0.0)
    kept = torch.where(gf <= torch.max(ga, gb), 1.0, 0.0)
    return q, (art * loss_term).sum() / den, (kept * loss_term).sum() / den


def nabf(a, b, f, L=1.5, modified=True):
    """calc_Nabf (metric.py:260-273)."""
    Qa, ga, gf = edge_preservation(a, f)
    Qb, gb, _ = edge_preservation(b, f)
    wa, wb = ga.pow(L), gb.pow(L)
    art = torch.where(gf > torch.max(ga, gb), 1.0, 0.0)
    if modified:
        return (art * ((1.0 - Qa) * wa + (1.0 - Qb) * wb)).sum() / (wa + wb).sum()
    return (art * ((2.0 - Qa - Qb) * (wa + wb))).sum() / (wa + wb).sum()


def labf(a, b, f, L=1.5):
    """calc_Labf (metric.py:277-286)."""
    Qa, ga, gf = edge_preservation(a, f)
    Qb, gb, _ = edge_preservation(b, f)
    wa, wb = ga.pow(L), gb.pow(L)
    kept = torch.where(gf <= torch.max(ga, gb), 1.0, 0.0)
    return (kept * ((1.0 - Qa) * wa + (1.0 - Qb) * wb)).sum() / (wa + wb).sum()


# ---- structural-similarity family -----------------------------------------------------
def metric_window(win_size=11, sigma=1.5):
    """metric.py:290-303 — sigma is explicit here (no window-size rule)."""
    col = gauss_taps(win_size, sigma).unsqueeze(1)
    return torch.mm(col, col.t())[None, None]


def ssim(a, b, win_size=11, data_range=255.0, use_padding=False, size_average=True, full=False):
    """metric.py:316-364 — global mean; window shrinks to min(win,H,W) with sigma 1.5."""
    k = min(win_size, a.shape[-2], a.shape[-1])
    w = metric_window(k).to(a)
    u, v = a.clone(), b.clone()
    mu_u, mu_v = blur(u, w, use_padding), blur(v, w, use_padding)
    uu, vv, uv = mu_u * mu_u, mu_v * mu_v, mu_u * mu_v
    var_u = (blur(u * u, w, use_padding) - uu).clamp(min=0)
    var_v = (blur(v * v, w, use_padding) - vv).clamp(min=0)
    cov = blur(u * v, w, use_padding) - uv
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    lum_n, lum_d = 2.0 * uv + c1, uu + vv + c1
    str_n, str_d = 2.0 * cov + c2, var_u + var_v + c2
    cs = str_n / str_d
    s = (lum_n * str_n) / (lum_d * str_d)
    if size_average:
        cs, s = cs.mean(), s.mean()
    return (s, cs) if full else s


def msssim(a, b, win_size=11, data_range=255.0, use_padding=False):
    """metric.py:368-402."""
    wts = torch.tensor(MS_WEIGHTS, dtype=torch.float32).to(a)
    u, v = a.clone(), b.clone()
    vals = []
    for lvl in range(len(wts)):
        s, cs = ssim(u, v, win_size, data_range, use_padding, full=True)
        if lvl < len(wts) - 1:
            vals.append(cs)
            u, v = halve(u), halve(v)
        else:
            vals.append(s)
    vals = torch.stack(vals, dim=0).clamp(min=1e-7)
    return torch.prod(vals ** wts, dim=0)


# ---- visual information fidelity ------------------------------------------------------
VIF_EPS = 1e-10                 # metric.py:407
VIF_NOISE = 0.005 * 255 * 255   # metric.py:408


def vif_maps(ref, dist, use_padding=False):
    """calc_vif (metric.py:406-458): per-scale VID, VIND and gain maps."""
    num, den, gain = [], [], []
    u, v = ref.clone(), dist.clone()
    for scale in range(1, 5):
        n = 2 ** (4 - scale + 1) + 1
        w = metric_window(n, n / 5).to(ref)
        if scale > 1:
            u = blur(u, w, use_padding)[..., ::2, ::2]
            v = blur(v, w, use_padding)[..., ::2, ::2]
        mu_u, mu_v = blur(u, w, use_padding), blur(v, w, use_padding)
        uu, vv, uv = mu_u * mu_u, mu_v * mu_v, mu_u * mu_v
        s1 = blur(u * u, w, use_padding) - uu
        s2 = blur(v * v, w, use_padding) - vv
        s12 = blur(u * v, w, use_padding) - uv
        s1[s1 < 0] = 0.0
        s2[s2 < 0] = 0.0
        g = s12 / (s1 + VIF_EPS)
        sv = s2 - g * s12
        low1 = s1 < VIF_EPS
        g[low1] = 0.0
        sv[low1] = s2[low1]
        s1[low1] = 0.0
        low2 = s2 < VIF_EPS
        g[low2] = 0.0
        sv[low2] = 0.0
        neg = g < 0
        sv[neg] = s2[neg]
        g[neg] = 0.0
        sv[sv < VIF_EPS] = VIF_EPS
        num.append(torch.log2(1 + g.pow(2) * s1 / (sv + VIF_NOISE)))
        den.append(torch.log2(1 + s1 / VIF_NOISE))
        gain.append(g)
    return num, den, gain


def viff(a, b, f, simple=True):
    """calc_viff (metric.py:461-491)."""
    n1, d1, g1 = vif_maps(a, f)
    n2, d2, g2 = vif_maps(b, f)
    if simple:
        sn1 = sn2 = sd1 = sd2 = 0.0
        for k in range(4):
            sn1 = sn1 + n1[k].sum()
            sn2 = sn2 + n2[k].sum()
            sd1 = sd1 + d1[k].sum()
            sd2 = sd2 + d2[k].sum()
        return sn1 / sd1 + sn2 / sd2
    p = torch.tensor([1.0, 0.0, 0.15, 1.0], dtype=torch.float32) / 2.15
    per_scale = torch.zeros(4)
    for k in range(4):
        pick = g1[k] < g2[k]
        per_scale[k] = torch.where(pick, n1[k], n2[k]).sum() / torch.where(pick, d1[k], d2[k]).sum()
    return (p * per_scale).sum()


# ---- eval.py composition --------------------------------------------------------------
def eval_pair(a, b, f):
    """The 16-metric row of eval.py:29-75 as a dict of python floats."""
    m = (mse(a, f) + mse(b, f)) * 0.5
    q, n, l = qabf(a, b, f, L=1.5, full=True)
    out = {
        'sd': std(f), 'ag': avg_gradient(f), 'sf': spatial_freq(f),
        'mse': m, 'psnr': psnr(m),
        'cc': (corrcoef(a, f) + corrcoef(b, f)) * 0.5, 'scd': scd(a, b, f),
        'en': entropy(f), 'ce': cross_entropy(a, f) + cross_entropy(b, f),
        'mi': mutual_info(a, f, normalized=True) + mutual_info(b, f, normalized=True),
        'qabf': q, 'nabf': n, 'labf': l,
        'ssim': (ssim(a, f) + ssim(b, f)) * 0.5,
        'msssim': (msssim(a, f) + msssim(b, f)) * 0.5,
        'viff': viff(a, b, f, simple=False),
    }
    return {k: out[k].item() for k in METRIC_NAMES}


def eval_pair_subset(a, b, f):
    """BASELINE config 4 subset: MS-SSIM + VIFF + Qabf/Nabf/Labf."""
    q, n, l = qabf(a, b, f, L=1.5, full=True)
    return {'qabf': q.item(), 'nabf': n.item(), 'labf': l.item(),
            'msssim': ((msssim(a, f) + msssim(b, f)) * 0.5).item(),
            'viff': viff(a, b, f, simple=False).item()}


def aggregate_columns(rows):
    """eval.py:231-266 — per metric: [mean, std, v0, v1, ...] where the std is taken
    over the list that already has the mean inserted at the front (reference quirk)."""
    cols = {}
    for name in METRIC_NAMES:
        vals = [r[name] for r in rows]
        vals.insert(0, np.mean(vals))
        vals.insert(1, np.std(vals))
        cols[name] = vals
    return cols
