"""Host logic of the drop-in metric functions for batches (N > 1), on CPU: `metric._combine_batch` turns the batched kernels'
per-pair rows into the ONE value the reference's function returns for the whole batch (metric.py:25-491 reduce over every
dimension).  Per-pair rows are produced here by the oracle, so no GPU is needed; the GPU test
test_batched_inputs_reduce_over_the_whole_batch checks the same rules end to end."""
import numpy as np
import torch

import mmif_b200  # noqa: F401
from mmif_b200 import _lib as L
from mmif_b200.core import metric as MM
from oracle import fusion_metric as OM
from oracle.fusion_loss import halve


def _batch():
    g = torch.Generator().manual_seed(77)
    a = torch.randint(0, 256, (3, 1, 96, 120), generator=g).double()
    b = torch.randint(0, 256, (3, 1, 96, 120), generator=g).double()
    a[1] = (a[1] * 0.3 + 100).floor()
    b[2] = (b[2] * 0.5).floor()
    return a, b, torch.floor((a + b) / 2)


def test_stats_rule():
    a, b, f = _batch()
    rows = torch.zeros(3, L.ST_COUNT, dtype=torch.float64)
    for i in range(3):
        x, y, z = a[i:i + 1], b[i:i + 1], f[i:i + 1]
        rows[i, :14] = torch.stack([z.mean(), OM.std(z), OM.avg_gradient(z), OM.spatial_freq(z), OM.mse(x, z), OM.mse(y, z),
                                    OM.corrcoef(x, z), OM.corrcoef(y, z), OM.scd(x, y, z), x.mean(), y.mean(), OM.std(x), OM.std(y),
                                    OM.corrcoef(x, y)])
    out = MM._combine_batch('stats', rows, None, None)[0]
    ref = [f.mean(), OM.std(f), OM.avg_gradient(f), OM.spatial_freq(f), OM.mse(a, f), OM.mse(b, f), OM.corrcoef(a, f),
           OM.corrcoef(b, f), OM.scd(a, b, f), a.mean(), b.mean(), OM.std(a), OM.std(b), OM.corrcoef(a, b)]
    np.testing.assert_allclose(out[:14].numpy(), [float(v) for v in ref], rtol=1e-11)
    assert abs(rows[:, 1].mean().item() - out[1].item()) > 1e-3 * out[1].item()      # not the mean of the per-pair values


def test_histogram_rule():
    a, b, f = _batch()
    counts = torch.zeros(3, L.HIST_WORDS, dtype=torch.int32)
    for i in range(3):
        x, y, z = a[i:i + 1].float(), b[i:i + 1].float(), f[i:i + 1].float()
        counts[i, 0:256], counts[i, 256:512], counts[i, 512:768] = (OM.hist_counts(t).int() for t in (x, y, z))
        counts[i, 768:768 + 65536] = torch.as_tensor(OM.joint_counts(x, z).numpy().astype(np.int32)).reshape(-1)
        counts[i, 768 + 65536:] = torch.as_tensor(OM.joint_counts(y, z).numpy().astype(np.int32)).reshape(-1)
    e = MM._combine_batch('hist', (counts, None), 1, f.numel())[0]
    ref = [OM.entropy(a), OM.entropy(b), OM.entropy(f), OM.joint_entropy(a, f), OM.joint_entropy(b, f), OM.cross_entropy(a, f),
           OM.cross_entropy(b, f), OM.mutual_info(a, f), OM.mutual_info(b, f), OM.mutual_info(a, f, True), OM.mutual_info(b, f, True)]
    np.testing.assert_allclose(e[:11].numpy(), [float(v) for v in ref], rtol=1e-12)


def test_ssim_msssim_viff_qabf_rules():
    a, b, f = _batch()
    rows = torch.zeros(3, L.MSSSIM_DOUBLES, dtype=torch.float64)
    for i in range(3):
        for off, src in ((0, a), (2, b)):
            u, v = src[i:i + 1].clone(), f[i:i + 1].clone()
            for lvl in range(5):
                s_, c_ = OM.ssim(u, v, full=True)
                rows[i, 2 + 4 * lvl + off], rows[i, 2 + 4 * lvl + off + 1] = s_, c_
                if lvl < 4:
                    u, v = halve(u), halve(v)
    m = MM._combine_batch('msssim', rows, None, None)[0]
    np.testing.assert_allclose([m[0].item(), m[1].item()], [OM.msssim(a, f).item(), OM.msssim(b, f).item()], rtol=1e-7)
    s4 = torch.stack([torch.stack([*OM.ssim(a[i:i + 1], f[i:i + 1], full=True), *OM.ssim(b[i:i + 1], f[i:i + 1], full=True)]) for i in range(3)])
    g = MM._combine_batch('ssim', s4, None, None)[0]
    np.testing.assert_allclose([g[0].item(), g[2].item()], [OM.ssim(a, f).item(), OM.ssim(b, f).item()], rtol=1e-12)
    rows = torch.zeros(3, L.VIFF_DOUBLES, dtype=torch.float64)
    for i in range(3):
        n1, d1, g1 = OM.vif_maps(a[i:i + 1], f[i:i + 1])
        n2, d2, g2 = OM.vif_maps(b[i:i + 1], f[i:i + 1])
        for k in range(4):
            pick = g1[k] < g2[k]
            rows[i, 2 + 6 * k:8 + 6 * k] = torch.stack([n1[k].sum(), d1[k].sum(), n2[k].sum(), d2[k].sum(),
                                                         torch.where(pick, n1[k], n2[k]).sum(), torch.where(pick, d1[k], d2[k]).sum()])
    v = MM._combine_batch('viff', rows, None, None)[0]
    np.testing.assert_allclose([v[0].item(), v[1].item()], [OM.viff(a, b, f, simple=False).item(), OM.viff(a, b, f, simple=True).item()], rtol=1e-6)
    rows = torch.zeros(3, 9, dtype=torch.float64)
    for i in range(3):
        Qa, ga, gf = OM.edge_preservation(a[i:i + 1], f[i:i + 1])
        Qb, gb, _ = OM.edge_preservation(b[i:i + 1], f[i:i + 1])
        wa, wb = ga.pow(1.5), gb.pow(1.5)
        loss = (1 - Qa) * wa + (1 - Qb) * wb
        art = gf > torch.max(ga, gb)
        rows[i, 4:9] = torch.stack([(Qa * wa + Qb * wb).sum(), (wa + wb).sum(), (loss * art).sum(), (loss * ~art).sum(),
                                    ((2 - Qa - Qb) * (wa + wb) * art).sum()])
    q = MM._combine_batch('qabf', rows, None, None)[0]
    rq, rn, rl = OM.qabf(a, b, f, L=1.5, full=True)
    np.testing.assert_allclose(q[:3].numpy(), [rq.item(), rn.item(), rl.item()], rtol=1e-12)
    np.testing.assert_allclose(q[3].item(), OM.nabf(a, b, f, modified=False).item(), rtol=1e-12)


def test_triple_tracker_learns_the_pair_of_sources():
    T = MM._Triples()
    a, b, f = (torch.rand(1, 1, 8, 8) for _ in range(3))
    assert T.of_pair(a, f) == (None, 0) and T.of_fused(f) is None
    tr, slot = T.of_pair(b, f)
    assert tr[0] is a and tr[1] is b and tr[2] is f and slot == 1
    assert T.of_pair(a, f)[1] == 0 and T.of_fused(f)[2] is f and T.of_fused(a) is None
    f.add_(1)                                           # an in-place change invalidates what was learned
    assert T.of_fused(f) is None and T.of_pair(a, f) == (None, 0)
