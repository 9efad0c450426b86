"""The single-pass loss+gradient kernel (fusion_loss_bwd_kernel<11, FAST, ZMODE=1>): the kernel SSIMLoss + PixelLoss +
GradLoss launch when imgf wants a gradient (train.py:64-71).  Parity of its loss block and of d(total)/d imgf against the
golden vectors of the real reference / the fp64 oracle, straight through the C ABI and through the drop-in modules, and
proof of WHICH kernel served each call (the library's launch counters and the CUPTI kernel names)."""
import ctypes

import numpy as np
import pytest
import torch

import cases
import gates
from oracle import fusion_loss as OL

pytestmark = pytest.mark.gpu
LG = np.load(cases.HERE + '/loss_golden.npz')
W3 = (1.0, 0.01, 0.1)          # train.py:302-308


def _mods():
    import mmif_b200  # noqa: F401
    from mmif_b200 import _lib as L
    from mmif_b200.core import loss as ML
    return L, ML


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x)).float()


def cabi_single_pass(a, b, f, w=W3, pixel=('max', 'l1'), grad=('max', 'l1')):
    """mmif_fusion_loss_fwd(want_grad=1) through ctypes -> (loss block as numpy doubles, float32 mirror, dF_unit tensor,
    handles for a following mmif_fusion_loss_bwd* call)."""
    L, ML = _mods()
    lib = L.load()
    A, B_, F_ = (t.cuda().contiguous() for t in (a, b, f))
    Bn, _, H, W = F_.shape
    L.ensure_device(F_.device)
    cfg = ML._cfg(1.0, pixel[0], grad[0], pixel[1], grad[1], *w)
    cfg.want_grad = 1
    nd = 4 + 6 * Bn
    assert lib.mmif_loss_out_doubles(Bn) == nd + (nd + 1) // 2
    out = torch.full((lib.mmif_loss_out_doubles(Bn),), float('nan'), dtype=torch.float64, device='cuda')
    ws = torch.zeros(lib.mmif_loss_workspace_bytes(Bn, H, W), dtype=torch.uint8, device='cuda')
    dU = torch.full_like(F_, float('nan'))
    before = L.launch_counts()
    L.check(lib.mmif_fusion_loss_fwd(A.data_ptr(), B_.data_ptr(), F_.data_ptr(), Bn, H, W, ctypes.byref(cfg), out.data_ptr(),
                                     dU.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr(F_.device)))
    after = L.launch_counts()
    assert after['loss_single_pass'] == before['loss_single_pass'] + 1 and after['loss_fwd'] == before['loss_fwd']
    torch.cuda.synchronize()
    blk = out[:nd].cpu().numpy()
    mirror = out[nd:].view(torch.float32)[:nd].cpu().numpy()
    return blk, mirror, dU, (lib, L, A, B_, F_, cfg, ws)


def check_block(name, blk, mirror, r32, r64, d32=None, d64=None):
    for k, nm in enumerate(('ssim', 'pixel', 'grad')):
        gates.assert_scalar(f'{name}/{nm}', blk[k], r32[k], r64[k])
    assert abs(blk[3] - (blk[0] + blk[1] + blk[2])) <= 1e-15 + 1e-12 * abs(blk[3])
    assert np.array_equal(mirror, blk.astype(np.float32)), 'the float32 mirror is the rounded double block'
    if d32 is not None:                     # per-sample ssim1, cs1, sigma1, ssim2, cs2, sigma2 (SSIM.forward dict)
        ps = blk[4:].reshape(-1, 6)
        for n in range(ps.shape[0]):
            for j in range(6):
                gates.assert_scalar(f'{name}/dict[{j},{n}]', ps[n, j], d32[j, n], d64[j, n])


def check_total_grad(name, got, a, b, f, per_term64, ref32_total=None, scale=1.0, w=W3):
    """d(total)/d imgf against the fp64 reference gradient.  Elements no L1 tie can touch: the max-norm gate (1e-5 of
    max|g|, or the reference's own fp32 error where that is larger).  Elements inside the counted tie neighbourhoods
    (gates.l1_tie_masks): the same gate widened by what flipped signs can move, 64 k_grad + 2 k_pixel."""
    tot64 = scale * per_term64.sum(axis=0)
    pix_mask, sob_mask, _ = gates.l1_tie_masks(a, b, f)
    mask = pix_mask | sob_mask
    gmax = np.abs(tot64).max()
    ref_err = 0.0
    if ref32_total is not None and (~mask).any():
        ref_err = (np.abs(scale * ref32_total - tot64) * ~mask).max() / gmax
    rtol = max(gates.RTOL, ref_err)
    frac, mx, masked = gates.masked_grad_report(got, tot64, mask, rtol=rtol)
    assert frac == 0.0, f'{name}: {frac:.2e} of the untied gradient elements beyond the gate, max {mx:.3e} (fp32 reference {ref_err:.3e}; {masked:.2e} ties)'
    assert masked <= (1.0 if name in cases.LOSS_GRAD_TIE_CASES else 1e-3), f'{name}: {masked:.2e} of the elements are ties'
    flip = abs(scale) * (64 * w[2] + 2 * w[1]) / f.size
    worst = (np.abs(got - tot64) * mask).max()
    assert worst <= rtol * gmax + flip * (1 + 1e-5), f'{name}: a tied element moved by {worst:.3e} > {rtol * gmax + flip:.3e}'


@pytest.mark.parametrize('name', cases.LOSS_CASES)
def test_cabi_single_pass_vs_golden(name):
    a, b, f = (T(x) for x in cases.loss_case(name))
    blk, mirror, dU, _ = cabi_single_pass(a, b, f)
    check_block(name, blk, mirror, LG[f'{name}/f32/loss'], LG[f'{name}/f64/loss'], LG[f'{name}/f32/ssim_dict'], LG[f'{name}/f64/ssim_dict'])
    got = dU.cpu().numpy()
    assert np.isfinite(got).all(), 'every element of dF_unit is written'
    check_total_grad(name, got, a.numpy(), b.numpy(), f.numpy(), LG[f'{name}/f64/grad'], LG[f'{name}/f32/grad_total'])


@pytest.mark.parametrize('shape,unbounded', [((1, 1024, 1224), False), ((8, 256, 256), True), ((2, 265, 546), False),
                                             ((1, 331, 371), True), ((3, 64, 1030), False)])
def test_cabi_single_pass_config_shapes_vs_live_oracle(shape, unbounded):
    """BASELINE configs[0] (1 x 1024 x 1224) and configs[1] (8 x 256 x 256 per rank with the network's UNBOUNDED imgf,
    model.py:177), two of the reference's sample sizes whose width is not a multiple of 4 (no TMA: the plain-load ring)
    and a ragged multi-strip width."""
    g = torch.Generator().manual_seed(sum(shape) + int(unbounded))
    a, b = (torch.rand((shape[0], 1) + shape[1:], generator=g) for _ in range(2))
    f = torch.randn((shape[0], 1) + shape[1:], generator=g) * 0.6 + 0.5 if unbounded else torch.rand((shape[0], 1) + shape[1:], generator=g)
    blk, mirror, dU, _ = cabi_single_pass(a, b, f)
    r32 = [t.item() for t in OL.train_objective(a, b, f)]
    l64, g64 = OL.train_objective_grad(a.double(), b.double(), f.double())
    check_block(str(shape), blk, mirror, r32, [t.item() for t in l64])
    f32 = f.clone().requires_grad_(True)
    sum(OL.train_objective(a, b, f32)).backward()
    check_total_grad(str(shape), dU.cpu().numpy(), a.numpy(), b.numpy(), f.numpy(), g64.numpy()[None], f32.grad.numpy())


@pytest.mark.parametrize('pixel,grad', [(('avg', 'l1'), ('avg', 'l1')), (('max', 'l2'), ('avg', 'l2')), (('avg', 'l2'), ('max', 'l1'))])
def test_cabi_single_pass_general_modes(pixel, grad):
    """The non-FAST instantiation (<11, 0, 1, 0>): avg / l2 combinations of the secondary terms."""
    a, b, f = (T(x) for x in cases.loss_case('rand_2x40x37'))
    blk, mirror, dU, _ = cabi_single_pass(a, b, f, pixel=pixel, grad=grad)
    ad, bd = a.double(), b.double()
    fd = f.double().requires_grad_(True)
    terms = (OL.ssim_loss(ad, bd, fd, 'ssim'), OL.pixel_loss(ad, bd, fd, pixel[1], 0.01, pixel[0]), OL.grad_loss(ad, bd, fd, grad[1], 0.1, grad[0]))
    r32 = (OL.ssim_loss(a, b, f, 'ssim').item(), OL.pixel_loss(a, b, f, pixel[1], 0.01, pixel[0]).item(), OL.grad_loss(a, b, f, grad[1], 0.1, grad[0]).item())
    check_block(f'{pixel}{grad}', blk, mirror, r32, [t.item() for t in terms])
    g64, = torch.autograd.grad(sum(terms), fd)
    frac, mx, where = gates.grad_report(dU.cpu().numpy(), g64.numpy())
    assert frac <= 1e-4, (frac, mx, where)


def test_cabi_backward_after_single_pass_rescale_and_recompute():
    """mmif_fusion_loss_bwd / _bwd3 with the single-pass buffer: equal upstream = rescale (in place and out of place, unit
    and non-unit), unequal upstream = recompute; NULL upstream pointers = 0."""
    name = 'rand_2x150x260'
    a, b, f = (T(x) for x in cases.loss_case(name))
    g64 = LG[f'{name}/f64/grad']
    blk, _, dU, (lib, L, A, B_, F_, cfg, ws) = cabi_single_pass(a, b, f)
    Bn, _, H, W = F_.shape
    st = L.stream_ptr(F_.device)
    unit = dU.clone()

    def bwd(gvals, unit_buf, dst, split=False):
        g = torch.tensor([0.0 if v is None else v for v in gvals], dtype=torch.float32, device='cuda')
        before = L.launch_counts()
        if split:
            ptr = [g[i:i + 1].data_ptr() if gvals[i] is not None else None for i in range(3)]
            L.check(lib.mmif_fusion_loss_bwd3(A.data_ptr(), B_.data_ptr(), F_.data_ptr(), Bn, H, W, ctypes.byref(cfg), ptr[0], ptr[1], ptr[2],
                                              unit_buf.data_ptr() if unit_buf is not None else None, dst.data_ptr(), ws.data_ptr(), ws.numel(), st))
        else:
            L.check(lib.mmif_fusion_loss_bwd(A.data_ptr(), B_.data_ptr(), F_.data_ptr(), Bn, H, W, ctypes.byref(cfg), g.data_ptr(),
                                             unit_buf.data_ptr() if unit_buf is not None else None, dst.data_ptr(), ws.data_ptr(), ws.numel(), st))
        torch.cuda.synchronize()
        after = L.launch_counts()
        return {k: after[k] - before[k] for k in after}

    # unit upstream, in place: the buffer already is the answer
    buf = unit.clone()
    d = bwd([1.0, 1.0, 1.0], buf, buf)
    assert d['rescale'] == 1 and torch.equal(buf, unit)
    # equal non-unit upstream, out of place and in place: a pure rescale of the single-pass buffer
    dst = torch.full_like(unit, float('nan'))
    bwd([2.5, 2.5, 2.5], unit, dst)
    assert torch.equal(dst, 2.5 * unit)
    buf = unit.clone()
    bwd([-0.75, -0.75, -0.75], buf, buf, split=True)
    assert torch.equal(buf, -0.75 * unit)
    # unequal upstream: recomputed (the buffer is ignored), equals the weighted sum of the fp64 per-term gradients
    dst = torch.full_like(unit, float('nan'))
    bwd([2.0, 3.0, -0.5], unit, dst)
    ref = 2.0 * g64[0] + 3.0 * g64[1] - 0.5 * g64[2]
    frac, mx, where = gates.grad_report(dst.cpu().numpy(), ref)
    assert frac <= 1e-4, (frac, mx, where)
    # split pointers with NULL = 0: only the SSIM term
    dst2 = torch.full_like(unit, float('nan'))
    bwd([1.0, None, None], None, dst2, split=True)
    frac, mx, where = gates.grad_report(dst2.cpu().numpy(), g64[0])
    assert frac == 0.0 or mx <= 2e-5, (frac, mx, where)
    # and the recomputed total equals the single-pass buffer to fp32 summation noise
    dst3 = torch.empty_like(unit)
    bwd([1.0, 1.0, 1.0], None, dst3)
    assert (dst3 - unit).abs().max().item() <= 2e-6 * unit.abs().max().item()


def _three(ML, A, B_, F_, w=W3):
    return (ML.SSIMLoss('ssim', weight=w[0])(A, B_, F_), ML.PixelLoss('l1', weight=w[1])(A, B_, F_, mode='max'),
            ML.GradLoss('l1', weight=w[2])(A, B_, F_, mode='max'))


@pytest.mark.parametrize('name', cases.LOSS_CASES)
def test_modules_total_backward_runs_the_single_pass_kernel(name):
    """train.py:64-71 through the drop-in modules: ONE single-pass launch serves the three modules, total.backward() is a
    rescale (no recompute), values and gradient match the real reference's golden vectors."""
    L, ML = _mods()
    a, b, f = (T(x) for x in cases.loss_case(name))
    A, B_, F_ = a.cuda(), b.cuda(), f.cuda().requires_grad_(True)
    c0 = L.launch_counts()
    l1, l2, l3 = _three(ML, A, B_, F_)
    c1 = L.launch_counts()
    assert c1['loss_single_pass'] - c0['loss_single_pass'] == 1, 'the three modules share ONE single-pass launch'
    assert c1['loss_fwd'] == c0['loss_fwd'] and c1['loss_bwd'] == c0['loss_bwd']
    node = l1.grad_fn
    assert node is l2.grad_fn and node is l3.grad_fn and node.single_pass and node.dF_unit is not None
    total = l1 + l2 + l3
    total.backward()
    torch.cuda.synchronize()
    c2 = L.launch_counts()
    assert c2['rescale'] - c1['rescale'] == 1 and node.dF_unit is None
    assert c2['loss_single_pass'] == c1['loss_single_pass'] and c2['loss_fwd'] == c1['loss_fwd']
    for k, (nm, v) in enumerate(zip(('ssim', 'pixel', 'grad'), (l1, l2, l3))):
        gates.assert_scalar(f'{name}/{nm}', v.item(), LG[f'{name}/f32/loss'][k], LG[f'{name}/f64/loss'][k])
    check_total_grad(name, F_.grad.cpu().numpy(), a.numpy(), b.numpy(), f.numpy(), LG[f'{name}/f64/grad'], LG[f'{name}/f32/grad_total'])


def test_modules_non_unit_unequal_and_retained_backward():
    L, ML = _mods()
    name = 'rand_3x64x96'
    a, b, f = (T(x) for x in cases.loss_case(name))
    g64 = LG[f'{name}/f64/grad']
    A, B_ = a.cuda(), b.cuda()
    # equal non-unit upstream: (2.5 * total).backward() -> rescale path
    F_ = f.cuda().requires_grad_(True)
    c0 = L.launch_counts()
    (2.5 * sum(_three(ML, A, B_, F_))).backward()
    c1 = L.launch_counts()
    assert c1['loss_single_pass'] - c0['loss_single_pass'] == 1 and c1['rescale'] - c0['rescale'] == 1
    frac, mx, where = gates.grad_report(F_.grad.cpu().numpy(), 2.5 * g64.sum(axis=0))
    assert frac <= 1e-4, (frac, mx, where)
    # unequal upstream: recompute
    F_ = f.cuda().requires_grad_(True)
    l1, l2, l3 = _three(ML, A, B_, F_)
    (2.0 * l1 + 3.0 * l2 - 0.5 * l3).backward()
    frac, mx, where = gates.grad_report(F_.grad.cpu().numpy(), 2.0 * g64[0] + 3.0 * g64[1] - 0.5 * g64[2])
    assert frac <= 1e-4, (frac, mx, where)
    # retain_graph: first backward consumes the single-pass buffer, the second recomputes — same numbers
    F_ = f.cuda().requires_grad_(True)
    total = sum(_three(ML, A, B_, F_))
    c0 = L.launch_counts()
    g1, = torch.autograd.grad(total, F_, retain_graph=True)
    g1 = g1.clone()
    g2, = torch.autograd.grad(total, F_)
    c1 = L.launch_counts()
    assert c1['loss_bwd'] - c0['loss_bwd'] >= 1, 'the second backward recomputes'
    assert c1['tmap_fail'] == 0, 'a tensor-map encoding failed (autograd backward thread without a bound context?)'
    assert (g1 - g2).abs().max().item() <= 2e-6 * g2.abs().max().item()
    frac, mx, where = gates.grad_report(g2.cpu().numpy(), g64.sum(axis=0))
    assert frac <= 1e-4, (frac, mx, where)


def test_no_grad_and_detached_inputs_run_the_forward_only_kernel():
    L, ML = _mods()
    a, b, f = (T(x).cuda() for x in cases.loss_case('rand_2x40x37'))
    c0 = L.launch_counts()
    v_plain = [t.item() for t in _three(ML, a, b, f)]                   # imgf does not require grad
    with torch.no_grad():
        v_nograd = [t.item() for t in _three(ML, a, b, f.clone().requires_grad_(True))]
    c1 = L.launch_counts()
    assert c1['loss_single_pass'] == c0['loss_single_pass'] and c1['loss_fwd'] - c0['loss_fwd'] == 2
    F_ = f.clone().requires_grad_(True)
    v_grad = [t.item() for t in _three(ML, a, b, F_)]
    assert L.launch_counts()['loss_single_pass'] - c1['loss_single_pass'] == 1
    for x, y, z in zip(v_plain, v_nograd, v_grad):
        assert x == y and abs(x - z) <= 2e-7 * abs(x)


def test_kernel_names_seen_by_cupti():
    """The same claim from the outside: the kernels CUPTI records for loss + backward through the modules."""
    L, ML = _mods()
    from torch.profiler import profile, ProfilerActivity
    a, b, f = (T(x).cuda() for x in cases.loss_case('rand_3x64x96'))
    F_ = f.clone().requires_grad_(True)
    sum(_three(ML, a, b, F_)).backward()          # warm-up outside the profile
    torch.cuda.synchronize()
    F_ = f.clone().requires_grad_(True)
    try:
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            sum(_three(ML, a, b, F_)).backward()
            torch.cuda.synchronize()
        names = [e.name for e in prof.events() if 'kernel' in e.name.lower()]
    except Exception as exc:  # pragma: no cover - CUPTI not available on the box
        pytest.skip(f'profiler unavailable: {exc}')
    if not names:
        pytest.skip('CUPTI recorded no kernels')
    # the single-pass instantiation: the warp-specialised kernel <11, FAST, ZMODE> (or, with MMIF_LOSS_WS=0, the 2-CTA kernel)
    norm = [n.replace('(bool)1', 'true').replace('(bool)0', 'false').replace('(bool)', '') for n in names]
    norm = [n.replace('(int)', '') for n in norm]
    assert any('fusion_loss_ws_kernel<11, true, 2>' in n or ('fusion_loss_ws_kernel' in n and '1, 2>' in n)
               or 'fusion_loss_bwd_kernel<11, true, true, false>' in n or ('fusion_loss_bwd_kernel' in n and '1, 1, 0' in n)
               for n in norm), names
    assert any('rescale_unit_kernel' in n for n in names), names
    assert not any('moment_fwd_kernel' in n for n in names), names


def test_train_step_loss_captures_into_a_cuda_graph():
    """The three module calls + total.backward() of train.py:64-71 are capturable (no host sync, no allocation outside the
    graph's pool, no host-side decision that depends on device data): replaying the graph on new data gives the eager numbers."""
    L, ML = _mods()
    g = torch.Generator().manual_seed(21)
    a0, b0, f0 = (torch.rand(4, 1, 96, 160, generator=g).cuda() for _ in range(3))
    a1, b1, f1 = (torch.rand(4, 1, 96, 160, generator=g).cuda() for _ in range(3))
    sa, sb = a0.clone(), b0.clone()
    sf = f0.clone().requires_grad_(True)
    mods = (ML.SSIMLoss('ssim', weight=1.0), ML.PixelLoss('l1', weight=0.01), ML.GradLoss('l1', weight=0.1))

    def step(A, B_, F_):
        tot = mods[0](A, B_, F_) + mods[1](A, B_, F_, mode='max') + mods[2](A, B_, F_, mode='max')
        tot.backward()
        return tot

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            sf.grad = None
            step(sa, sb, sf)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    sf.grad = None
    with torch.cuda.graph(graph):
        stot = step(sa, sb, sf)
    for A, B_, F_ in ((a1, b1, f1), (a0, b0, f0)):
        sa.copy_(A); sb.copy_(B_)
        with torch.no_grad():
            sf.copy_(F_)
        graph.replay()
        torch.cuda.synchronize()
        Fe = F_.clone().requires_grad_(True)
        etot = step(A, B_, Fe)
        assert stot.item() == etot.item()
        assert torch.equal(sf.grad, Fe.grad)


def test_calling_the_losses_twice_on_the_same_tensors_builds_two_graphs():
    """The reference builds a new autograd graph on every call; the one-launch memo must not make a second
    loss(...).backward() on the same tensors fail (or silently reuse a consumed graph)."""
    L, ML = _mods()
    a, b, f = (T(x).cuda() for x in cases.loss_case('rand_2x40x37'))
    F_ = f.clone().requires_grad_(True)
    sum(_three(ML, a, b, F_)).backward()
    g1 = F_.grad.clone()
    F_.grad = None
    c0 = L.launch_counts()
    sum(_three(ML, a, b, F_)).backward()            # same tensor objects, same versions
    assert L.launch_counts()['loss_single_pass'] - c0['loss_single_pass'] == 1
    assert torch.equal(F_.grad, g1)


def test_fastcall_binding_equals_the_ctypes_binding():
    """The per-step entries through the CPython fast-call binding and through ctypes: the same kernels, bit-equal outputs."""
    L, ML = _mods()
    fc = L.fastcall()
    assert fc is not None
    lib = L.load()
    a, b, f = (T(x).cuda() for x in cases.loss_case('rand_3x64x96'))
    B, H, W = a.shape[0], a.shape[2], a.shape[3]
    cfg = ML._cfg(1.0, 'max', 'max', 'l1', 'l1', 1.0, 0.01, 0.1)
    cfg.want_grad = 1
    st = L.stream_int(a.device)
    res = []
    for fast in (False, True):
        out = torch.zeros(lib.mmif_loss_out_doubles(B), dtype=torch.float64, device='cuda')
        ws = torch.zeros(lib.mmif_loss_workspace_bytes(B, H, W), dtype=torch.uint8, device='cuda')
        dU = torch.zeros_like(f)
        if fast:
            rc = fc.loss_fwd(a.data_ptr(), b.data_ptr(), f.data_ptr(), B, H, W, ctypes.addressof(cfg), out.data_ptr(), dU.data_ptr(),
                             ws.data_ptr(), ws.numel(), st)
        else:
            rc = lib.mmif_fusion_loss_fwd(a.data_ptr(), b.data_ptr(), f.data_ptr(), B, H, W, ctypes.byref(cfg), out.data_ptr(),
                                          dU.data_ptr(), ws.data_ptr(), ws.numel(), st)
        assert rc == 0
        g = torch.tensor([2.0, 0.5, 3.0], device='cuda')
        dF = torch.zeros_like(f)
        cfg0 = ML._cfg(1.0, 'max', 'max', 'l1', 'l1', 1.0, 0.01, 0.1)
        if fast:
            rc = fc.loss_bwd3(a.data_ptr(), b.data_ptr(), f.data_ptr(), B, H, W, ctypes.addressof(cfg0), g[0:1].data_ptr(),
                              g[1:2].data_ptr(), g[2:3].data_ptr(), None, dF.data_ptr(), ws.data_ptr(), ws.numel(), st)
        else:
            rc = lib.mmif_fusion_loss_bwd3(a.data_ptr(), b.data_ptr(), f.data_ptr(), B, H, W, ctypes.byref(cfg0), g[0:1].data_ptr(),
                                           g[1:2].data_ptr(), g[2:3].data_ptr(), None, dF.data_ptr(), ws.data_ptr(), ws.numel(), st)
        assert rc == 0
        torch.cuda.synchronize()
        res.append((out.clone(), dU.clone(), dF.clone()))
    for x, y in zip(*res):
        assert torch.equal(x, y)


def test_warp_specialised_kernel_equals_the_two_cta_kernel_on_a_shape_sweep(monkeypatch):
    """The training configuration runs on fusion_loss_ws_kernel (one CTA per SM, four warp groups over named barriers);
    MMIF_LOSS_WS=0 puts it back on the 2-CTA kernel.  Same arithmetic: on 40 shapes from 11 x 11 up (single batch / single
    strip, ragged last batches, widths TMA cannot describe, several strips and segments) the loss blocks agree to 1e-6 and
    the gradients to 2e-6 of max|g| (they are bit-equal when both kernels cut the rows alike; different row segments move
    the tile shift constants), for the single-pass launch and for the recomputing backward with unequal upstreams."""
    L, ML = _mods()
    lib = L.load()
    rng = np.random.RandomState(11)
    shapes = [(1, 11, 11), (2, 11, 40), (1, 40, 11), (1, 12, 128), (3, 19, 109), (1, 27, 217), (2, 64, 96), (1, 300, 2050)]
    shapes += [(int(rng.randint(1, 4)), int(rng.randint(11, 90)), int(rng.randint(11, 330))) for _ in range(32)]
    for si, (B, H, W) in enumerate(shapes):
        g = torch.Generator().manual_seed(B * 1000003 + H * 1009 + W)
        a, b, f = (torch.rand(B, 1, H, W, generator=g).cuda() for _ in range(3))
        st = L.stream_int(a.device)
        res = {}
        # every third shape in the avg / l2 modes (the general instantiation: warp-specialised up to 32 Mpix per launch)
        modes = ('avg', 'avg', 'l2', 'l2') if si % 3 == 2 else ('max', 'max', 'l1', 'l1')
        for ws_on in ('0', '1'):
            monkeypatch.setenv('MMIF_LOSS_WS', ws_on)
            cfg = ML._cfg(1.0, *modes, 1.0, 0.01, 0.1)
            cfg.want_grad = 1
            out = torch.zeros(lib.mmif_loss_out_doubles(B), dtype=torch.float64, device='cuda')
            ws = torch.zeros(lib.mmif_loss_workspace_bytes(B, H, W), dtype=torch.uint8, device='cuda')
            dU = torch.full_like(f, float('nan'))
            L.check(lib.mmif_fusion_loss_fwd(a.data_ptr(), b.data_ptr(), f.data_ptr(), B, H, W, ctypes.byref(cfg), out.data_ptr(),
                                             dU.data_ptr(), ws.data_ptr(), ws.numel(), st))
            cfg0 = ML._cfg(1.0, *modes, 1.0, 0.01, 0.1)
            up = torch.tensor([1.5, -0.25, 3.0], device='cuda')
            dF = torch.full_like(f, float('nan'))
            L.check(lib.mmif_fusion_loss_bwd(a.data_ptr(), b.data_ptr(), f.data_ptr(), B, H, W, ctypes.byref(cfg0), up.data_ptr(), None,
                                             dF.data_ptr(), ws.data_ptr(), ws.numel(), st))
            torch.cuda.synchronize()
            res[ws_on] = (out[:L.LOSS_HEAD + L.LOSS_PER_SAMPLE * B].clone(), dU.clone(), dF.clone())
        (o0, u0, d0), (o1, u1, d1) = res['0'], res['1']
        assert torch.isfinite(u1).all() and torch.isfinite(d1).all(), (B, H, W)
        rel = ((o0 - o1).abs() / o0.abs().clamp_min(1e-12)).max().item()
        assert rel <= 1e-6, ((B, H, W), rel)
        for x0, x1, nm in ((u0, u1, 'single-pass'), (d0, d1, 'recomputing')):
            err = (x0 - x1).abs().max().item() / max(x0.abs().max().item(), 1e-30)
            assert err <= 2e-6, ((B, H, W), nm, err)


def test_cabi_want_grad_2_writes_the_loss_values_and_the_same_gradient():
    """want_grad = 2 (what the drop-in modules pass: the training step reads the three loss values only): the same gradient
    bit for bit, the same loss head and per-sample SSIM means as want_grad = 1, cs / sigma entries written as 0."""
    L, ML = _mods()
    lib = L.load()
    for name in ('rand_3x64x96', 'ir_crop_max'):
        a, b, f = (T(x).cuda() for x in cases.loss_case(name))
        B, H, W = a.shape[0], a.shape[2], a.shape[3]
        st = L.stream_int(a.device)
        res = {}
        for wg in (1, 2):
            cfg = ML._cfg(1.0, 'max', 'max', 'l1', 'l1', 1.0, 0.01, 0.1)
            cfg.want_grad = wg
            out = torch.full((lib.mmif_loss_out_doubles(B),), float('nan'), dtype=torch.float64, device='cuda')
            ws = torch.zeros(lib.mmif_loss_workspace_bytes(B, H, W), dtype=torch.uint8, device='cuda')
            dU = torch.full_like(f, float('nan'))
            L.check(lib.mmif_fusion_loss_fwd(a.data_ptr(), b.data_ptr(), f.data_ptr(), B, H, W, ctypes.byref(cfg), out.data_ptr(),
                                             dU.data_ptr(), ws.data_ptr(), ws.numel(), st))
            torch.cuda.synchronize()
            res[wg] = (out[:L.LOSS_HEAD + L.LOSS_PER_SAMPLE * B].clone(), dU.clone())
        (o1, g1), (o2, g2) = res[1], res[2]
        assert torch.equal(g1, g2)
        assert torch.equal(o1[:L.LOSS_HEAD], o2[:L.LOSS_HEAD])
        p1 = o1[L.LOSS_HEAD:].view(B, L.LOSS_PER_SAMPLE)
        p2 = o2[L.LOSS_HEAD:].view(B, L.LOSS_PER_SAMPLE)
        assert torch.equal(p1[:, [0, 3]], p2[:, [0, 3]])                    # ssim means of the two pairs
        assert (p2[:, [1, 2, 4, 5]] == 0).all() and (p1[:, [1, 2, 4, 5]] != 0).all()
