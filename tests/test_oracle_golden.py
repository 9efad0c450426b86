"""Pins the CPU oracle (``oracle/``) to outputs of the real reference (``tests/golden``)."""
import numpy as np
import pytest
import torch

import cases
from oracle import fusion_loss as OL, fusion_metric as OM

LG = np.load(cases.HERE + '/loss_golden.npz')
MG = np.load(cases.HERE + '/metric_golden.npz')


def T(x, dt=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dt)


@pytest.mark.parametrize('name', cases.LOSS_CASES)
@pytest.mark.parametrize('tag', ['f32', 'f64'])
def test_loss_terms_and_grad(name, tag):
    dt = torch.float32 if tag == 'f32' else torch.float64
    a, b, f = (T(x, dt) for x in cases.loss_case(name))
    (l1, l2, l3), _ = OL.train_objective_grad(a, b, f)
    got = np.array([l1.item(), l2.item(), l3.item()])
    np.testing.assert_allclose(got, LG[f'{name}/{tag}/loss'], rtol=1e-6 if tag == 'f32' else 1e-12, atol=0)
    if tag == 'f64':
        ref = LG[f'{name}/f64/grad']
        for k, up in enumerate(((1, 0, 0), (0, 1, 0), (0, 0, 1))):
            _, g = OL.train_objective_grad(a, b, f, upstream=up)
            np.testing.assert_allclose(g.numpy(), ref[k], rtol=1e-10, atol=1e-16)


@pytest.mark.parametrize('name', cases.LOSS_CASES)
def test_ssim_dict_and_secondary_modes(name):
    a, b, f = (T(x) for x in cases.loss_case(name))
    w = OL.window2d(11, 1.5)
    d1 = OL.ssim(a, f, window=w, data_range=1.0)
    d2 = OL.ssim(b, f, window=w, data_range=1.0)
    got = np.stack([d1['ssim'].numpy(), d1['cs'].numpy(), d1['sigma'].numpy(),
                    d2['ssim'].numpy(), d2['cs'].numpy(), d2['sigma'].numpy()]).astype(np.float64)
    np.testing.assert_allclose(got, LG[f'{name}/f32/ssim_dict'], rtol=1e-6)
    sec = [OL.pixel_loss(a, b, f, 'l1', 0.01, 'avg').item(), OL.pixel_loss(a, b, f, 'l2', 0.01, 'max').item(),
           OL.grad_loss(a, b, f, 'l1', 0.1, 'avg').item(), OL.grad_loss(a, b, f, 'l2', 0.1, 'max').item(),
           OL.ssim_loss(a, b, f, 'w-ssim').item(), OL.tv_loss(f - a).item()]
    np.testing.assert_allclose(sec, LG[f'{name}/f32/secondary'], rtol=1e-6)


def test_loss_smoke_main_anchor():
    """loss.py:388-423 smoke values (4-decimal anchors in SURVEY.md §4) through the oracle."""
    torch.manual_seed(0)
    x1, x2, y = torch.rand(2, 1, 256, 256), torch.rand(2, 1, 256, 256), torch.rand(2, 1, 256, 256)
    chk = np.array([x1.double().sum().item(), x2.double().sum().item(), y.double().sum().item()])
    if not np.allclose(chk, LG['smoke_main/inputs_checksum'], rtol=0, atol=1e-6):
        pytest.skip('torch CPU generator stream differs from the one the golden was made with')
    got = [OL.ssim_loss(x1, x2, y, 'ssim').item(), OL.pixel_loss(x1, x2, y, 'l1', 0.01).item(),
           OL.grad_loss(x1, x2, y, 'l1', 0.1).item(), OL.tv_loss(y - x1).item()]
    np.testing.assert_allclose(got, LG['smoke_main/loss'], rtol=1e-6)
    assert [round(v, 4) for v in got] == [0.9945, 0.0033, 0.0927, 0.9310]


@pytest.mark.parametrize('name', cases.METRIC_CASES)
@pytest.mark.parametrize('tag', ['f32', 'f64'])
def test_metric_suite(name, tag):
    dt = torch.float32 if tag == 'f32' else torch.float64
    a, b, f = (T(x, dt) for x in cases.metric_case(name))
    r = OM.eval_pair(a, b, f)
    got = np.array([r[k] for k in OM.METRIC_NAMES])
    np.testing.assert_allclose(got, MG[f'{name}/{tag}/metrics'], rtol=1e-6 if tag == 'f32' else 1e-12, atol=1e-9)
    extra = [OM.mean(f).item(), OM.nabf(a, b, f, modified=False).item(), OM.viff(a, b, f, simple=True).item(),
             OM.mutual_info(a, f).item(), OM.ssim(a / 255.0, f / 255.0, data_range=1.0).item(),
             OM.psnr(OM.mse(a, f), root=True).item()]
    np.testing.assert_allclose(extra, MG[f'{name}/{tag}/extra'], rtol=1e-6 if tag == 'f32' else 1e-12, atol=1e-9)


def _dense(idx, cnt):
    j = np.zeros(65536, np.int64)
    j[idx] = cnt
    return j.reshape(256, 256)


@pytest.mark.parametrize('name', cases.METRIC_CASES)
def test_histograms_bit_exact(name):
    a, b, f = (T(x) for x in cases.metric_case(name))
    for key, img in (('hist_a', a), ('hist_b', b), ('hist_f', f)):
        assert np.array_equal(OM.hist_counts(img).to(torch.int64).numpy(), MG[f'{name}/{key}'])
    assert np.array_equal(OM.joint_counts(a, f).numpy().astype(np.int64),
                          _dense(MG[f'{name}/joint_af_idx'], MG[f'{name}/joint_af_cnt']))
    assert np.array_equal(OM.joint_counts(b, f).numpy().astype(np.int64),
                          _dense(MG[f'{name}/joint_bf_idx'], MG[f'{name}/joint_bf_cnt']))


def test_histogram_edge_rule():
    v, w = (T(x) for x in cases.hist_edge_vector())
    assert np.array_equal(OM.hist_counts(v).to(torch.int64).numpy(), MG['hist_edge/hist_v'])
    assert np.array_equal(OM.hist_counts(w).to(torch.int64).numpy(), MG['hist_edge/hist_w'])
    assert np.array_equal(OM.joint_counts(v, w).numpy().astype(np.int64),
                          _dense(MG['hist_edge/joint_idx'], MG['hist_edge/joint_cnt']))


def test_metric_smoke_main_anchor():
    """metric.py:494-551 smoke values (SURVEY.md §4 anchors)."""
    torch.manual_seed(0)
    x1, x2, y = (torch.rand(1, 1, 256, 256) * 255.0 for _ in range(3))
    chk = np.array([x1.double().sum().item(), x2.double().sum().item(), y.double().sum().item()])
    if not np.allclose(chk, MG['smoke_main/inputs_checksum'], rtol=0, atol=1e-3):
        pytest.skip('torch CPU generator stream differs from the one the golden was made with')
    r = OM.eval_pair(x1, x2, y)
    np.testing.assert_allclose([r[k] for k in OM.METRIC_NAMES], MG['smoke_main/metrics'], rtol=1e-6, atol=1e-9)
    anchors = dict(sd=73.7165, ag=93.4801, sf=147.3422, mse=0.1666, psnr=7.7832, cc=0.0008, scd=-0.0048,
                   en=7.9915, ce=0.0112, mi=0.2040, qabf=0.2562, nabf=0.1585, labf=0.5854, ssim=0.0060,
                   msssim=0.0906, viff=0.0163)
    for k, v in anchors.items():
        assert abs(r[k] - v) < 6e-5, (k, r[k], v)


def test_aggregate_quirk():
    rows = [{k: float(i + j) for j, k in enumerate(OM.METRIC_NAMES)} for i in range(4)]
    cols = OM.aggregate_columns(rows)
    vals = [0.0, 1.0, 2.0, 3.0]
    m = np.mean(vals)
    assert cols['sd'][0] == m and cols['sd'][1] == np.std([m] + vals) and cols['sd'][2:] == vals


WG = np.load(cases.HERE + '/ssim_windows_golden.npz')


@pytest.mark.parametrize('win', [9, 7, 5, 3])
@pytest.mark.parametrize('pad', [False, True])
def test_ssim_other_windows_pinned_to_the_reference(win, pad):
    """SSIM(win_size=k) (loss.py:163-185; window sigma 0.15 (k-1), loss.py:34): oracle vs the real reference."""
    a, _, f = (T(x) for x in cases.loss_case('rand_3x64x96'))
    for tag, dt, rtol in (('f32', torch.float32, 1e-6), ('f64', torch.float64, 1e-12)):
        d = OL.ssim(a.to(dt), f.to(dt), win, None, 1.0, pad)
        got = np.stack([d[k].numpy().astype(np.float64) for k in ('ssim', 'cs', 'sigma')])
        np.testing.assert_allclose(got, WG[f'ssim/win{win}/pad{int(pad)}/{tag}'], rtol=rtol, atol=0)


@pytest.mark.parametrize('win,pad', [(7, False), (5, True)])
def test_msssim_other_windows_pinned_to_the_reference(win, pad):
    x, y = T(WG['ms/x']), T(WG['ms/y'])
    for tag, dt, rtol in (('f32', torch.float32, 1e-6), ('f64', torch.float64, 1e-12)):
        got = OL.msssim(x.to(dt), y.to(dt), win, None, None, 1.0, pad).numpy().astype(np.float64)
        np.testing.assert_allclose(got, WG[f'msssim/win{win}/pad{int(pad)}/{tag}'], rtol=rtol, atol=0)
