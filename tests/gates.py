"""Parity gates of SURVEY.md §8(c), shared by the GPU tests."""
import numpy as np

RTOL = 1e-5      # north_star: floating-point results within 1e-5 relative in fp32
ATOL = 1e-7      # absolute floor for scalars that are ~0 (cc, scd, nabf on random data)
BAND = 1.0       # width of the accepted band around ref64, in units of |ref32 - ref64|: "at least as close to the fp64
                 # truth as the reference's own fp32 evaluation is" 


def scalar_ok(new, ref32, ref64):
    """|new-ref32| <= RTOL|ref32| (+ATOL)  OR  new at least as close to the fp64 truth as the fp32
    reference is (the reference's own fp32 noise reaches ~8e-6 on real images)."""
    new, ref32, ref64 = float(new), float(ref32), float(ref64)
    if not np.isfinite(new):
        return np.isnan(new) and np.isnan(ref32)
    if abs(new - ref32) <= RTOL * abs(ref32) + ATOL:
        return True
    # Ill-conditioned quantity (the reference's own fp32 and fp64 evaluations disagree by more than
    # the tolerance, e.g. VIFF's gain-based source selection on flat regions): stay within the
    # reference's own uncertainty band around the fp64 value.
    return abs(new - ref64) <= BAND * abs(ref32 - ref64)


def assert_scalar(name, new, ref32, ref64, allowance=None):
    """`allowance`: optional callable -> absolute slack for a quantity with a DISCRETE ill-conditioning the band cannot see
    (a hard mask decided by a comparison of two rounded values); evaluated only when the plain gate fails, and logged."""
    if allowance is not None and not scalar_ok(new, ref32, ref64):
        slack = float(allowance())
        BAND_LOG.append((name + f' [tie allowance {slack:.3e}]', abs(float(new) - float(ref64)) / max(slack, 1e-300)))
        assert abs(float(new) - float(ref64)) <= abs(float(ref32) - float(ref64)) + slack, (
            f'{name}: new={float(new)!r} ref32={float(ref32)!r} ref64={float(ref64)!r} tie allowance {slack:.3e}')
        return
    if abs(float(new) - float(ref32)) > RTOL * abs(float(ref32)) + ATOL and float(ref32) != float(ref64):
        BAND_LOG.append((name, abs(float(new) - float(ref64)) / abs(float(ref32) - float(ref64))))
    assert scalar_ok(new, ref32, ref64), (
        f'{name}: new={float(new)!r} ref32={float(ref32)!r} ref64={float(ref64)!r} '
        f'rel32={abs(float(new) - float(ref32)) / max(abs(float(ref32)), 1e-300):.3e} '
        f'rel64={abs(float(new) - float(ref64)) / max(abs(float(ref64)), 1e-300):.3e}')


def grad_report(new, ref64, rtol=RTOL):
    """max-norm gate against the fp64 oracle; returns (ok_fraction_bad, max_rel, where)."""
    new = np.asarray(new, dtype=np.float64)
    ref64 = np.asarray(ref64, dtype=np.float64)
    scale = np.abs(ref64).max()
    diff = np.abs(new - ref64)
    bad = diff > rtol * scale
    where = np.unravel_index(np.argmax(diff), diff.shape)
    return bad.mean(), (diff.max() / scale if scale > 0 else diff.max()), where


# ---- L1 sign ties of the pixel / Sobel terms ("excluding elements where the sign() argument is within fp32
# rounding of zero (count them)") ------------------------------------------------------------------------------------
BAND_LOG = []        # (name, |new-ref64| / |ref32-ref64|) of every scalar that passed through the band, for the report


def l1_tie_masks(a, b, f, tol=2e-6):
    """float64 evaluation of the arguments whose sign() enters d/d imgf of PixelLoss('l1', mode='max') and
    GradLoss('l1', mode='max') (loss.py:294-304, 330-344).  Returns (pixel_mask, sobel_mask, pixel_exact_zero):
    boolean (B,1,H,W) arrays of the gradient ELEMENTS that a tie can change, and the positions where the pixel argument
    is EXACTLY zero.
      * pixel term: d = imgf - max(img1, img2) is ONE subtraction, exact zeros are the same in every precision
        (sign(0) = 0: the gradient there must be exactly 0), so only NEAR ties 0 < |d| <= tol are masked;
      * Sobel term: D = S(imgf) - max(S(img1), S(img2)) and the responses gx, gy of imgf are sums of 6-8 rounded terms.
        On 8-bit data (k/255) they are integer combinations that are exactly 0 in exact arithmetic at a large share of
        the positions and +-1e-8 after fp32 rounding, so |arg| <= tol INCLUDING 0 is a tie (the reference's own fp32 and
        fp64 gradients differ on 25 % of the elements of `ir_crop_max`); the 3x3 neighbourhood of every tied position
        is masked (the Sobel adjoint spreads a sign over it).  gx on the first / last column and gy on the
        first / last row are structural zeros of the reflect padding (not ties)."""
    import torch
    import torch.nn.functional as F
    a, b, f = (torch.as_tensor(x, dtype=torch.float64) for x in (a, b, f))
    kx = torch.tensor([[-1., 0., 1.], [-2., 0., 2.], [-1., 0., 1.]], dtype=torch.float64).view(1, 1, 3, 3)
    ky = torch.tensor([[-1., -2., -1.], [0., 0., 0.], [1., 2., 1.]], dtype=torch.float64).view(1, 1, 3, 3)

    def sob(u):
        p = F.pad(u, (1, 1, 1, 1), mode='reflect')
        gx, gy = F.conv2d(p, kx), F.conv2d(p, ky)
        return gx, gy, gx.abs() + gy.abs()

    d = f - torch.maximum(a, b)
    pix_mask = (d.abs() <= tol) & (d != 0)
    gxf, gyf, sf = sob(f)
    _, _, s1 = sob(a)
    _, _, s2 = sob(b)
    D = sf - torch.maximum(s1, s2)
    # the reflect padding makes gx of the first / last column and gy of the first / last row a difference of IDENTICAL values:
    # exactly 0 in every precision, a structural zero and not a tie
    tx, ty = gxf.abs() <= tol, gyf.abs() <= tol
    tx[..., :, 0] = False
    tx[..., :, -1] = False
    ty[..., 0, :] = False
    ty[..., -1, :] = False
    tD = D.abs() <= tol
    for r in (0, -1):                   # the four corners: gx = gy = 0 structurally for every image, so D = 0 exactly
        for c in (0, -1):
            tD[..., r, c] = False
    pos = tD | tx | ty
    sob_mask = F.max_pool2d(pos.double(), 3, 1, 1) > 0
    return pix_mask.numpy(), sob_mask.numpy(), (d == 0).numpy()


def masked_grad_report(new, ref64, mask, rtol=RTOL):
    """grad_report over the elements NOT in `mask`; returns (bad_fraction_of_unmasked, max_rel, masked_fraction)."""
    new = np.asarray(new, dtype=np.float64)
    ref64 = np.asarray(ref64, dtype=np.float64)
    keep = ~np.asarray(mask, dtype=bool)
    scale = np.abs(ref64).max()
    diff = np.abs(new - ref64) * keep
    bad = diff > rtol * scale
    return bad.sum() / max(int(keep.sum()), 1), (diff.max() / scale if scale > 0 else diff.max()), 1.0 - keep.mean()


def qabf_tie_allowance(a, b, f, L=1.5, tol=2e-6):
    """Nabf / Labf (metric.py:233-286) put each pixel's loss weight into one of two sums by the hard comparison
    g_f > max(g_a, g_b) of rounded fp32 edge strengths.  A pixel whose two sides agree to `tol` relative can land on either
    side in any fp32 evaluation (the reference's included); this is the total share of such pixels, the amount by which
    nabf and labf may legitimately move (their sum does not)."""
    import torch
    from oracle import fusion_metric as OM
    a, b, f = (torch.as_tensor(x, dtype=torch.float64) for x in (a, b, f))
    Qa, ga, gf = OM.edge_preservation(a, f)
    Qb, gb, _ = OM.edge_preservation(b, f)
    wa, wb = ga.pow(L), gb.pow(L)
    loss = (1.0 - Qa) * wa + (1.0 - Qb) * wb
    gmax = torch.max(ga, gb)
    near = (gf - gmax).abs() <= tol * gmax.clamp(min=1e-30)
    return ((loss * near).sum() / (wa + wb).sum()).item()


def viff_tie_allowance(a, b, f, noise=1e-6, gtol=1e-5):
    """calc_viff(simple=False) (metric.py:461-491) takes, per pixel and scale, the (numerator, denominator) of the source with
    the smaller gain: pick = g1 < g2, a HARD selection between two values that can differ a lot.  Where a local variance is
    at the fp32 rounding floor of blur(x^2) - mu^2 (|s| <= noise * E[x^2]: saturated / flat patches — the reference's own
    fp32 variance there is rounding noise around 0 that the `< 1e-10` rules of metric.py:436-452 then branch on) or the two
    gains agree to gtol, the selection is decided by rounding in ANY fp32 evaluation; the reference's fp32 and fp64 per-scale
    ratios differ by 1e-4..8e-4 on such images while their weighted sum happens to agree to 1e-5.  Returns the amount by
    which the weighted VIFF can move if every such pixel flips (first-order bound), evaluated in float64."""
    import torch
    from oracle import fusion_metric as OM
    from oracle.fusion_loss import blur
    a, b, f = (torch.as_tensor(x, dtype=torch.float64) for x in (a, b, f))

    def var_maps(ref, dist):
        out, u, v = [], ref.clone(), dist.clone()
        for scale in range(1, 5):
            n = 2 ** (4 - scale + 1) + 1
            w = OM.metric_window(n, n / 5).to(ref)
            if scale > 1:
                u, v = blur(u, w)[..., ::2, ::2], blur(v, w)[..., ::2, ::2]
            mu_u, mu_v = blur(u, w), blur(v, w)
            e1, e2 = blur(u * u, w), blur(v * v, w)
            out.append((e1 - mu_u * mu_u, e2 - mu_v * mu_v, e1, e2))
        return out

    n1, d1, g1 = OM.vif_maps(a, f)
    n2, d2, g2 = OM.vif_maps(b, f)
    m1, m2 = var_maps(a, f), var_maps(b, f)
    p = [1.0 / 2.15, 0.0, 0.15 / 2.15, 1.0 / 2.15]
    total = 0.0
    for k in range(4):
        pick = g1[k] < g2[k]
        num = torch.where(pick, n1[k], n2[k]).sum()
        den = torch.where(pick, d1[k], d2[k]).sum()
        s1a, s2f, e1a, e2f = m1[k]
        s1b, _, e1b, _ = m2[k]
        unc = (s1a <= noise * e1a) | (s1b <= noise * e1b) | (s2f <= noise * e2f) | \
              ((g1[k] - g2[k]).abs() <= gtol * torch.maximum(g1[k].abs(), g2[k].abs()))
        dn, dd = ((n1[k] - n2[k]).abs() * unc).sum(), ((d1[k] - d2[k]).abs() * unc).sum()
        total += p[k] * (dn / den + num / den ** 2 * dd).item()
    return total
