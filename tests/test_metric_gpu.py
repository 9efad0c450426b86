"""GPU parity of the metric suite against the golden vectors of the real reference and the live
oracle.  Histogram counts are compared bit-exactly; floating-point metrics with the dual gate of
SURVEY.md 8(c) (gates.py)."""
import numpy as np
import pytest
import torch

import cases
import gates
from oracle import fusion_metric as OM

pytestmark = pytest.mark.gpu
MG = np.load(cases.HERE + '/metric_golden.npz')


def _mods():
    import mmif_b200  # noqa: F401
    from mmif_b200.core import metric as MM
    return MM


def T(x, dt=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dt)


def _dense(idx, cnt):
    j = np.zeros(65536, np.int64)
    j[idx] = cnt
    return j.reshape(256, 256)


def ref_row_via_dropins(MM, a, b, f):
    """eval.py:29-75 written against the drop-in functions exactly as the reference script does."""
    sd, ag, sf = MM.calc_std(f), MM.calc_ag(f), MM.calc_sf(f)
    mse = (MM.calc_mse(a, f) + MM.calc_mse(b, f)) * 0.5
    psnr = MM.calc_psnr(mse)
    cc = (MM.calc_cc(a, f) + MM.calc_cc(b, f)) * 0.5
    scd = MM.calc_scd(a, b, f)
    en = MM.calc_entropy(f)
    ce = MM.calc_cross_ent(a, f) + MM.calc_cross_ent(b, f)
    mi = MM.calc_mul_info(a, f, normalized=True) + MM.calc_mul_info(b, f, normalized=True)
    qabf, nabf, labf = MM.calc_Qabf(a, b, f, L=1.5, full=True)
    ssim = (MM.calc_ssim(a, f) + MM.calc_ssim(b, f)) * 0.5
    msssim = (MM.calc_msssim(a, f) + MM.calc_msssim(b, f)) * 0.5
    viff = MM.calc_viff(a, b, f, simple=False)
    assert mi.dtype == torch.float64 and sd.dtype == torch.float32 and sd.dim() == 0
    vals = [sd, ag, sf, mse, psnr, cc, scd, en, ce, mi, qabf, nabf, labf, ssim, msssim, viff]
    return [v.item() for v in vals]


@pytest.mark.parametrize('name', cases.METRIC_CASES)
@pytest.mark.parametrize('where', ['cuda', 'cpu'])
def test_dropin_functions_vs_golden(name, where):
    """CPU inputs are what eval.py really passes (eval.py:198-200); CUDA inputs what test.py passes."""
    MM = _mods()
    a, b, f = (T(x).to(where) for x in cases.metric_case(name))
    got = ref_row_via_dropins(MM, a, b, f)
    r32, r64 = MG[f'{name}/f32/metrics'], MG[f'{name}/f64/metrics']
    for k, nm in enumerate(OM.METRIC_NAMES):
        gates.assert_scalar(f'{name}/{nm}', got[k], r32[k], r64[k])


@pytest.mark.parametrize('name', cases.METRIC_CASES)
def test_fused_suite_vs_golden_and_extras(name):
    MM = _mods()
    a, b, f = (T(x).cuda() for x in cases.metric_case(name))
    row = MM.eval_metrics(a, b, f)
    r32, r64 = MG[f'{name}/f32/metrics'], MG[f'{name}/f64/metrics']
    for k, nm in enumerate(OM.METRIC_NAMES):
        gates.assert_scalar(f'{name}/suite/{nm}', row[nm], r32[k], r64[k])
    extra = [MM.calc_mean(f).item(), MM.calc_Nabf(a, b, f, modified=False).item(), MM.calc_viff(a, b, f, simple=True).item(),
             MM.calc_mul_info(a, f).item(), MM.calc_ssim(a / 255.0, f / 255.0, data_range=1.0).item(),
             MM.calc_psnr(MM.calc_mse(a, f), root=True).item()]
    e32, e64 = MG[f'{name}/f32/extra'], MG[f'{name}/f64/extra']
    for k, nm in enumerate(('mean', 'nabf_unmodified', 'viff_simple', 'mi_raw', 'ssim_L1', 'psnr_root')):
        gates.assert_scalar(f'{name}/{nm}', extra[k], e32[k], e64[k])


@pytest.mark.parametrize('name', cases.METRIC_CASES)
def test_histograms_bit_exact(name):
    MM = _mods()
    a, b, f = (T(x).cuda() for x in cases.metric_case(name))
    ha, hb, hf, jaf, jbf = MM.histograms(a, b, f)
    assert np.array_equal(ha.numpy(), MG[f'{name}/hist_a'])
    assert np.array_equal(hb.numpy(), MG[f'{name}/hist_b'])
    assert np.array_equal(hf.numpy(), MG[f'{name}/hist_f'])
    assert np.array_equal(jaf.numpy(), _dense(MG[f'{name}/joint_af_idx'], MG[f'{name}/joint_af_cnt']))
    assert np.array_equal(jbf.numpy(), _dense(MG[f'{name}/joint_bf_idx'], MG[f'{name}/joint_bf_cnt']))


def test_histogram_edge_rule_bit_exact():
    """value 256.0 -> bin 255, -0.0 -> bin 0, <0 / >256 / NaN / inf dropped (pair dropped if either is)."""
    MM = _mods()
    v, w = (T(x).cuda() for x in cases.hist_edge_vector())
    hv, _, hw, jvw, _ = MM.histograms(v, v, w)
    assert np.array_equal(hv.numpy(), MG['hist_edge/hist_v'])
    assert np.array_equal(hw.numpy(), MG['hist_edge/hist_w'])
    assert np.array_equal(jvw.numpy(), _dense(MG['hist_edge/joint_idx'], MG['hist_edge/joint_cnt']))


@pytest.mark.parametrize('shape', [(480, 640), (1024, 1224)])
def test_config_shapes_vs_live_oracle(shape):
    """BASELINE configs 3 / 4 shapes (640x480 TNO, 1224x1024 polarization), integer-valued images."""
    MM = _mods()
    g = torch.Generator().manual_seed(shape[0])
    a = torch.randint(0, 256, (1, 1) + shape, generator=g).float()
    b = torch.randint(0, 256, (1, 1) + shape, generator=g).float()
    f = torch.floor((a + b) / 2)
    r32 = OM.eval_pair(a, b, f)
    r64 = OM.eval_pair(a.double(), b.double(), f.double())
    row = MM.eval_metrics(a.cuda(), b.cuda(), f.cuda())
    for nm in OM.METRIC_NAMES:
        gates.assert_scalar(f'{shape}/{nm}', row[nm], r32[nm], r64[nm])
    ha, hb, hf, jaf, jbf = MM.histograms(a.cuda(), b.cuda(), f.cuda())
    assert np.array_equal(hf.numpy(), OM.hist_counts(f).to(torch.int64).numpy())
    assert np.array_equal(jaf.numpy(), OM.joint_counts(a, f).numpy().astype(np.int64))
    assert np.array_equal(jbf.numpy(), OM.joint_counts(b, f).numpy().astype(np.int64))
    assert int(ha.sum()) == a.numel() and int(jaf.sum()) == a.numel()


def test_batched_suite_equals_per_pair_and_is_deterministic():
    MM = _mods()
    g = torch.Generator().manual_seed(11)
    a = torch.randint(0, 256, (5, 1, 120, 200), generator=g).float().cuda()
    b = torch.randint(0, 256, (5, 1, 120, 200), generator=g).float().cuda()
    f = torch.maximum(a, b)
    rows = MM.eval_metrics_batch(a, b, f)
    rows2 = MM.eval_metrics_batch(a, b, f)
    assert torch.equal(rows, rows2)
    for n in range(5):
        one = MM.eval_metrics_batch(a[n:n + 1], b[n:n + 1], f[n:n + 1])[0]
        np.testing.assert_allclose(rows[n].cpu().numpy(), one.cpu().numpy(), rtol=1e-7, atol=1e-12)  # row blocking differs with N


def test_constant_images_and_errors():
    """sigma = 0 paths: cc is 0/0 = nan in the reference too; Qabf 0/0 -> 0 weights; VIF eps branches."""
    MM = _mods()
    a = torch.full((1, 1, 64, 64), 100.0)
    b = torch.full((1, 1, 64, 64), 50.0)
    f = torch.full((1, 1, 64, 64), 75.0)
    r32 = OM.eval_pair(a, b, f)
    r64 = OM.eval_pair(a.double(), b.double(), f.double())
    row = MM.eval_metrics(a.cuda(), b.cuda(), f.cuda())
    for nm in ('sd', 'ag', 'sf', 'mse', 'psnr', 'en', 'ce', 'ssim', 'msssim'):
        gates.assert_scalar(f'const/{nm}', row[nm], r32[nm], r64[nm])
    for nm in ('cc', 'scd', 'qabf', 'nabf', 'labf'):
        assert np.isnan(row[nm]) == np.isnan(r32[nm]), (nm, row[nm], r32[nm])
    with pytest.raises(Exception):
        MM.calc_viff(torch.rand(1, 1, 20, 20).cuda(), torch.rand(1, 1, 20, 20).cuda(), torch.rand(1, 1, 20, 20).cuda())


@pytest.mark.parametrize('n_pairs', [1, 16])
def test_histogram_counter_overflow_and_split_modes(n_pairs):
    """Bins far above the 15-bit shared-memory counters (flat / two-valued images), in the split mode
    (few pairs: several CTAs add into one global histogram) and in the fused mode (>= 16 pairs: one CTA
    per joint finishes from shared memory); entropies against the live oracle."""
    MM = _mods()
    g = torch.Generator().manual_seed(5)
    H, W = 300, 400
    a = torch.full((n_pairs, 1, H, W), 255.0)
    a[:, :, :, 100:] = torch.randint(0, 2, (n_pairs, 1, H, W - 100), generator=g).float() * 7.0
    b = torch.full((n_pairs, 1, H, W), 256.0)                  # lands in bin 255
    b[:, :, 17, :] = -1.0                                       # dropped samples: marginal of f only
    f = torch.full((n_pairs, 1, H, W), 3.0)
    f[:, :, :50, :] = torch.randint(0, 256, (n_pairs, 1, 50, W), generator=g).float()
    counts, ent = MM.hist_raw(a.cuda(), b.cuda(), f.cuda())
    counts = counts.to(torch.int64).cpu().numpy()
    ent = ent.cpu().numpy()
    for n in range(n_pairs):
        an, bn, fn = a[n:n + 1], b[n:n + 1], f[n:n + 1]
        assert np.array_equal(counts[n, 0:256], OM.hist_counts(an).to(torch.int64).numpy())
        assert np.array_equal(counts[n, 256:512], OM.hist_counts(bn).to(torch.int64).numpy())
        assert np.array_equal(counts[n, 512:768], OM.hist_counts(fn).to(torch.int64).numpy())
        assert np.array_equal(counts[n, 768:768 + 65536].reshape(256, 256), OM.joint_counts(an, fn).numpy().astype(np.int64))
        assert np.array_equal(counts[n, 768 + 65536:].reshape(256, 256), OM.joint_counts(bn, fn).numpy().astype(np.int64))
        for k, (x, y) in enumerate(((an, fn), (bn, fn))):
            r32 = OM.mutual_info(x, y, normalized=True).item()
            r64 = OM.mutual_info(x.double(), y.double(), normalized=True).item()
            gates.assert_scalar(f'ovf/nmi{k}', ent[n, 9 + k], r32, r64)
            gates.assert_scalar(f'ovf/ce{k}', ent[n, 5 + k], OM.cross_entropy(x, y).item(), OM.cross_entropy(x.double(), y.double()).item())
        gates.assert_scalar('ovf/en_f', ent[n, 2], OM.entropy(fn).item(), OM.entropy(fn.double()).item())


def test_use_padding_and_maps_vs_oracle():
    """calc_ssim / calc_msssim with use_padding=True (reflect pad per level) and size_average=False (maps)."""
    MM = _mods()
    g = torch.Generator().manual_seed(21)
    a = torch.randint(0, 256, (1, 1, 90, 117), generator=g).float()
    f = torch.floor(0.5 * (a + torch.randint(0, 256, (1, 1, 90, 117), generator=g).float()))
    for fn_new, fn_ref in ((MM.calc_ssim, OM.ssim), (MM.calc_msssim, OM.msssim)):
        got = fn_new(a.cuda(), f.cuda(), use_padding=True).item()
        gates.assert_scalar(fn_ref.__name__ + '/pad', got, fn_ref(a, f, use_padding=True).item(),
                            fn_ref(a.double(), f.double(), use_padding=True).item())
    for pad in (False, True):
        s_map, cs_map = MM.calc_ssim(a.cuda(), f.cuda(), use_padding=pad, size_average=False, full=True)
        r_s, r_cs = OM.ssim(a.double(), f.double(), use_padding=pad, size_average=False, full=True)
        assert tuple(s_map.shape) == tuple(r_s.shape)
        np.testing.assert_allclose(s_map.cpu().numpy(), r_s.numpy(), rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(cs_map.cpu().numpy(), r_cs.numpy(), rtol=2e-5, atol=2e-6)


def test_uint8_ingest_equals_float_path():
    MM = _mods()
    g = torch.Generator().manual_seed(4)
    a = torch.randint(0, 256, (3, 1, 131, 203), generator=g, dtype=torch.uint8)
    b = torch.randint(0, 256, (3, 1, 131, 203), generator=g, dtype=torch.uint8)
    f = torch.maximum(a, b)
    ref = MM.eval_metrics_batch(a.float().cuda(), b.float().cuda(), f.float().cuda())
    for src in ((a, b, f), (a.cuda(), b.cuda(), f.cuda()), (a.pin_memory(), b.pin_memory(), f.pin_memory())):
        got = MM.eval_metrics_batch_u8(*src)
        assert torch.equal(got, ref)


def test_pipelined_host_ingest_equals_batched_entry():
    """eval_metrics_batch_host: chunked upload overlapped with the suite; rows = the one-launch batched entry up to the
    fp32 summation order of a different batch split (histogram metrics EN/MI/CE exactly)."""
    MM = _mods()
    g = torch.Generator().manual_seed(14)
    a = torch.randint(0, 256, (9, 1, 140, 212), generator=g, dtype=torch.uint8)
    b = torch.randint(0, 256, (9, 1, 140, 212), generator=g, dtype=torch.uint8)
    f = ((a.int() + b.int()) // 2).to(torch.uint8)
    ref = MM.eval_metrics_batch(a.float().cuda(), b.float().cuda(), f.float().cuda()).cpu().numpy()
    for src in ((a.pin_memory(), b.pin_memory(), f.pin_memory()), (a, b, f),
                (a.float().pin_memory(), b.float().pin_memory(), f.float().pin_memory())):
        for chunks in (4, 1, 16):
            got = MM.eval_metrics_batch_host(*src, chunks=chunks).cpu().numpy()
            np.testing.assert_allclose(got, ref, rtol=5e-6, atol=2e-8)
    names = list(OM.METRIC_NAMES)
    got = MM.eval_metrics_batch_host(a.pin_memory(), b.pin_memory(), f.pin_memory()).cpu().numpy()
    for nm in ('en', 'mi', 'ce'):
        if nm in names:
            assert np.array_equal(got[:, names.index(nm)], ref[:, names.index(nm)]), nm
    with pytest.raises(Exception):
        MM.eval_metrics_batch_host(a.cuda(), b.cuda(), f.cuda())


def test_test_py_post_step_one_pass():
    """test.py:49-73: avg SSIM (data_range=1.0) and the denorm() image; the image must be bit-exact."""
    MM = _mods()
    g = torch.Generator().manual_seed(12)
    a, b = (torch.rand(2, 1, 75, 133, generator=g) for _ in range(2))
    f = (0.5 * (a + b) + 0.4 * (torch.rand(2, 1, 75, 133, generator=g) - 0.5)) * 1.3 - 0.1      # leaves [0, 1]: the clip matters
    f[0, 0, 3, 5] = float('nan'); f[1, 0, 0, 0] = 1.0; f[1, 0, 74, 132] = 0.999999
    f_ok = torch.nan_to_num(f, nan=0.5)
    ssim, img8 = MM.test_post_step(a.cuda(), b.cuda(), f_ok.cuda())
    for n in range(2):
        r32 = ((OM.ssim(a[n:n + 1], f_ok[n:n + 1], data_range=1.0) + OM.ssim(b[n:n + 1], f_ok[n:n + 1], data_range=1.0)) * 0.5).item()
        r64 = ((OM.ssim(a[n:n + 1].double(), f_ok[n:n + 1].double(), data_range=1.0) +
                OM.ssim(b[n:n + 1].double(), f_ok[n:n + 1].double(), data_range=1.0)) * 0.5).item()
        gates.assert_scalar('test_post/ssim', ssim[n].item(), r32, r64)
        ref8 = (f_ok[n].numpy().clip(0, 1) * 255.0).transpose((1, 2, 0)).astype(np.uint8)[..., 0]      # data/transform.py:32-35
        assert np.array_equal(img8[n].cpu().numpy(), ref8)
    _, img8n = MM.test_post_step(a.cuda(), b.cuda(), f.cuda())
    assert int(img8n[0, 3, 5]) == 0 and np.array_equal(img8n[1].cpu().numpy(), img8[1].cpu().numpy())


def _eval_py_row(MM, a, b, f):
    """eval.py:29-75 verbatim call sequence on the drop-in functions."""
    sd, ag, sf = MM.calc_std(f), MM.calc_ag(f), MM.calc_sf(f)
    mse = (MM.calc_mse(a, f) + MM.calc_mse(b, f)) * 0.5
    psnr = MM.calc_psnr(mse)
    cc = (MM.calc_cc(a, f) + MM.calc_cc(b, f)) * 0.5
    scd = MM.calc_scd(a, b, f)
    en = MM.calc_entropy(f)
    ce = MM.calc_cross_ent(a, f) + MM.calc_cross_ent(b, f)
    mi = MM.calc_mul_info(a, f, normalized=True) + MM.calc_mul_info(b, f, normalized=True)
    q, n, l = MM.calc_Qabf(a, b, f, L=1.5, full=True)
    ssim = (MM.calc_ssim(a, f) + MM.calc_ssim(b, f)) * 0.5
    ms = (MM.calc_msssim(a, f) + MM.calc_msssim(b, f)) * 0.5
    viff = MM.calc_viff(a, b, f, simple=False)
    return [v.item() for v in (sd, ag, sf, mse, psnr, cc, scd, en, ce, mi, q, n, l, ssim, ms, viff)]


@pytest.mark.parametrize('where', ['cpu', 'cuda'])
def test_eval_py_call_sequence_shares_launches_across_the_triple(where):
    """The one- and two-image calls of eval.py:29-75 are served by three-image launches once the triple (a, b, f) is known
    (learned from calc_mse(a, f), calc_mse(b, f)): one launch per kernel family on the triple, the same numbers as the
    batched suite and as the same functions called in isolation (each on its own argument list)."""
    import mmif_b200  # noqa: F401
    from mmif_b200 import _lib as L
    MM = _mods()
    g = torch.Generator().manual_seed(123)
    a = torch.randint(0, 256, (1, 1, 150, 210), generator=g).float()
    b = torch.randint(0, 256, (1, 1, 150, 210), generator=g).float()
    f = torch.floor((a + b) / 2)
    if where == 'cuda':
        a, b, f = a.cuda(), b.cuda(), f.cuda()
    c0 = L.launch_counts()
    row = _eval_py_row(MM, a, b, f)
    c1 = L.launch_counts()
    # pixel kernel: (f,f,f), (a,a,f), (a,b,f), qabf = 4; hist = 1; then ssim 1+1, ms-ssim 5+4+1, viff 4+3+1 kernels
    assert c1['metric'] - c0['metric'] <= 4 + 1 + 1 + 5 + 4 + 2, (c0, c1)
    assert c1['moment_fwd'] - c0['moment_fwd'] <= 1 + 5 + 4, (c0, c1)
    suite = MM.eval_metrics_batch(a.cuda(), b.cuda(), f.cuda())[0].cpu().numpy()
    np.testing.assert_allclose(row, suite, rtol=2e-6, atol=1e-9)
    # the same functions on fresh tensors in an order that never reveals the triple: identical values
    a2, b2, f2 = a.clone(), b.clone(), f.clone()
    iso = [MM.calc_cc(b2, f2).item(), MM.calc_cross_ent(b2, f2).item(), MM.calc_ssim(b2.clone(), f2).item()]
    shared = [MM.calc_cc(b, f).item(), MM.calc_cross_ent(b, f).item(), MM.calc_ssim(b, f).item()]
    np.testing.assert_allclose(iso, shared, rtol=1e-6, atol=1e-12)
    # an in-place change of the fused image invalidates everything that was learned
    f.add_(1.0)
    assert abs(MM.calc_mse(a, f).item() - row[3]) > 0


@pytest.mark.parametrize('where', ['cuda', 'cpu'])
def test_batched_inputs_reduce_over_the_whole_batch(where):
    """Every reference metric reduces over ALL dimensions (metric.py:25-491): for N > 1 the drop-in functions return the
    statistic of the whole batch as one population — global std / cc / scd, one histogram, ratios of batch-wide sums, per-level
    global means in MS-SSIM — not a mean of per-pair values."""
    MM = _mods()
    g = torch.Generator().manual_seed(77)
    a = torch.randint(0, 256, (3, 1, 96, 120), generator=g).float()
    b = torch.randint(0, 256, (3, 1, 96, 120), generator=g).float()
    a[1] = (a[1] * 0.3 + 100).floor()            # different means / contrasts per pair: global != mean of per-pair
    b[2] = (b[2] * 0.5).floor()
    f = torch.floor((a + b) / 2)
    A, B_, F_ = (t.cuda() for t in (a, b, f)) if where == 'cuda' else (a, b, f)
    got = {'mean': MM.calc_mean(F_), 'sd': MM.calc_std(F_), 'ag': MM.calc_ag(F_), 'sf': MM.calc_sf(F_), 'mse': MM.calc_mse(A, F_),
           'cc': MM.calc_cc(B_, F_), 'scd': MM.calc_scd(A, B_, F_), 'en': MM.calc_entropy(F_), 'ce': MM.calc_cross_ent(A, F_),
           'mi': MM.calc_mul_info(B_, F_), 'nmi': MM.calc_mul_info(A, F_, normalized=True), 'ssim': MM.calc_ssim(A, F_),
           'msssim': MM.calc_msssim(B_, F_), 'viff': MM.calc_viff(A, B_, F_, simple=False), 'viff_s': MM.calc_viff(A, B_, F_, simple=True)}
    q, n, l = MM.calc_Qabf(A, B_, F_, L=1.5, full=True)
    got.update(qabf=q, nabf=n, labf=l)

    def ref(dt):
        x, y, z = a.to(dt), b.to(dt), f.to(dt)
        r = {'mean': OM.mean(z), 'sd': OM.std(z), 'ag': OM.avg_gradient(z), 'sf': OM.spatial_freq(z), 'mse': OM.mse(x, z),
             'cc': OM.corrcoef(y, z), 'scd': OM.scd(x, y, z), 'en': OM.entropy(z), 'ce': OM.cross_entropy(x, z),
             'mi': OM.mutual_info(y, z), 'nmi': OM.mutual_info(x, z, normalized=True), 'ssim': OM.ssim(x, z),
             'msssim': OM.msssim(y, z), 'viff': OM.viff(x, y, z, simple=False), 'viff_s': OM.viff(x, y, z, simple=True)}
        r['qabf'], r['nabf'], r['labf'] = OM.qabf(x, y, z, L=1.5, full=True)
        return {k: float(v) for k, v in r.items()}

    r32, r64 = ref(torch.float32), ref(torch.float64)
    for k, v in got.items():
        assert v.dim() == 0 and v.is_cuda == (where == 'cuda')
        gates.assert_scalar(f'batched/{k}', v.item(), r32[k], r64[k])
    # and it is NOT the mean of the per-pair values where the two differ
    per_pair = np.mean([MM.calc_std(F_[i:i + 1]).item() for i in range(3)])
    assert abs(per_pair - got['sd'].item()) > 1e-3 * got['sd'].item()
    ha, hb, hf, jaf, jbf = MM.histograms(A, B_, F_)
    assert int(hf.sum()) == f.numel() and np.array_equal(hf.numpy(), OM.hist_counts(f).to(torch.int64).numpy())
