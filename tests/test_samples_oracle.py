"""Pins the CPU oracle to the REAL reference on the reference's own sample images (tests/golden/samples_golden.npz, written
by make_golden_samples.py from /root/reference), and — where the byte-compiled reference artefact oracle/_ref is present —
checks the oracle against the reference itself, live.  A bounded subset so the CPU suite stays within minutes; the GPU
tests (test_samples_gpu.py) cover all 21 pairs x 3 fused images."""
import numpy as np
import pytest
import torch

import samples as S
from oracle import fusion_loss as OL, fusion_metric as OM

SG = np.load(S.HERE + '/samples_golden.npz')
# the five pairs whose width is not a multiple of 4 (no TMA on the GPU), one 480x640 pair, one full polar pair
SUBSET = [('infrared/037.png', 'avg_noise'), ('infrared/049.png', 'max'), ('infrared/100.png', 'avg_round'),
          ('infrared/108.png', 'avg_noise'), ('infrared/175.png', 'max'), ('infrared/00537D.png', 'avg_round'),
          ('polar/1.jpg', 'avg_noise')]


def T(x, dt=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dt)


def test_fixture_inventory():
    names = S.names()
    assert len(names) == 21 and len([n for n in names if n.startswith('polar/')]) == 5
    shapes = {n: S.pair(n)[0].shape for n in names}
    assert shapes['polar/1.jpg'] == (1024, 1224) and shapes['infrared/037.png'] == (265, 546)
    assert sorted(n for n, s in shapes.items() if s[1] % 4) == ['infrared/037.png', 'infrared/049.png', 'infrared/100.png',
                                                                 'infrared/108.png', 'infrared/175.png']
    for n in names:
        a, b = S.pair(n)
        assert a.dtype == np.uint8 and a.shape == b.shape
        for kind in S.KINDS:
            assert f'{n}/{kind}/f32/metrics' in SG.files and f'{n}/{kind}/f64/loss' in SG.files


@pytest.mark.parametrize('name,kind', SUBSET)
def test_oracle_metrics_on_real_images(name, kind):
    a, b, f = (T(x) for x in S.case(name, kind))
    r = OM.eval_pair(a, b, f)
    got = np.array([r[k] for k in OM.METRIC_NAMES])
    np.testing.assert_allclose(got, SG[f'{name}/{kind}/f32/metrics'], rtol=1e-6, atol=1e-9)
    hist = np.stack([OM.hist_counts(x).to(torch.int64).numpy() for x in (a, b, f)])
    assert np.array_equal(hist, SG[f'{name}/{kind}/hist'])
    assert np.array_equal(S.joint_checksum(OM.joint_counts(a, f).numpy()), SG[f'{name}/{kind}/joint_af'])
    assert np.array_equal(S.joint_checksum(OM.joint_counts(b, f).numpy()), SG[f'{name}/{kind}/joint_bf'])


@pytest.mark.parametrize('name,kind', SUBSET[:6])
def test_oracle_loss_and_gradient_on_real_images(name, kind):
    a, b, f = S.case(name, kind)
    au, bu, fu = (S.unit(x) for x in (a, b, f))
    (l1, l2, l3), _ = OL.train_objective_grad(T(au), T(bu), T(fu))
    np.testing.assert_allclose([l1.item(), l2.item(), l3.item()], SG[f'{name}/{kind}/f32/loss'], rtol=1e-6)
    (l1, l2, l3), g = OL.train_objective_grad(T(au, torch.float64), T(bu, torch.float64), T(fu, torch.float64))
    np.testing.assert_allclose([l1.item(), l2.item(), l3.item()], SG[f'{name}/{kind}/f64/loss'], rtol=1e-12)
    probe = S.grad_probe(fu.shape, S.names().index(name))
    np.testing.assert_allclose(g.numpy().reshape(-1)[probe], SG[f'{name}/{kind}/f64/grad_probe'][3], rtol=1e-9, atol=1e-18)
    np.testing.assert_allclose(np.abs(g.numpy()).sum(), SG[f'{name}/{kind}/f64/grad_l1'][3], rtol=1e-9)


def test_densefuse_output_case():
    """imgf = the reference's DenseFuse(seed 0) output: an UNBOUNDED fused image on a real pair."""
    name = 'infrared/05.png'
    a, b = (x.astype(np.float32)[None, None] for x in S.pair(name))
    fu = SG[f'{name}/densefuse/imgf']
    (l1, l2, l3), _ = OL.train_objective_grad(T(S.unit(a)), T(S.unit(b)), T(S.unit((fu * np.float32(255.0)).astype(np.float32))))
    np.testing.assert_allclose([l1.item(), l2.item(), l3.item()], SG[f'{name}/densefuse/f32/loss'], rtol=1e-6)


def test_oracle_equals_the_built_reference_live():
    """oracle/ vs the reference's own modules (oracle/_ref, byte-compiled from /root/reference by oracle/build_ref.py)."""
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip('oracle/_ref not built on this machine')
    RL, RM, _ = build_ref.load()
    g = torch.Generator().manual_seed(11)
    a, b, f = (torch.rand(2, 1, 70, 93, generator=g) for _ in range(3))
    y = f.clone().requires_grad_(True)
    terms = (RL.SSIMLoss('ssim', weight=1.0)(a, b, y), RL.PixelLoss('l1', weight=0.01)(a, b, y, mode='max'),
             RL.GradLoss('l1', weight=0.1)(a, b, y, mode='max'))
    sum(terms).backward()
    (l1, l2, l3), go = OL.train_objective_grad(a, b, f)
    assert [t.item() for t in terms] == [l1.item(), l2.item(), l3.item()]
    assert torch.equal(y.grad, go)
    x, z, w = (torch.randint(0, 256, (1, 1, 64, 80), generator=g).float() for _ in range(3))
    r = OM.eval_pair(x, z, w)
    assert r['ssim'] == ((RM.calc_ssim(x, w) + RM.calc_ssim(z, w)) * 0.5).item()
    assert r['viff'] == RM.calc_viff(x, z, w, simple=False).item()
    q, n, l = RM.calc_Qabf(x, z, w, L=1.5, full=True)
    assert (r['qabf'], r['nabf'], r['labf']) == (q.item(), n.item(), l.item())
    assert r['mi'] == (RM.calc_mul_info(x, w, normalized=True) + RM.calc_mul_info(z, w, normalized=True)).item()
