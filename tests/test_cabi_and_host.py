"""CPU-only checks: the C-ABI library loads and exports every symbol include/mmif_b200.h declares,
argument validation works without a GPU, the drop-in surface mirrors the reference's names and
error behaviour, and the multi-process plumbing (gloo, world_size 2) gives single-process results."""
import ctypes
import os
import re
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'mmif_b200.h')
LIBP = os.path.join(ROOT, 'multi-modal-image-fusion_b200', 'libmmif_b200.so')


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(mmif_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIBP), 'build the library first: make (or __graft_entry__.build())'
    lib = ctypes.CDLL(LIBP)
    names = declared_functions()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f'{n} declared in mmif_b200.h but not exported'
    lib.mmif_version.restype = ctypes.c_int
    assert lib.mmif_version() == 100


def test_python_binding_lists_the_same_symbols():
    import mmif_b200
    from mmif_b200 import _lib
    assert sorted(_lib.SYMBOLS) == declared_functions()
    _lib.load()


def test_argument_validation_without_gpu():
    import mmif_b200  # noqa: F401
    from mmif_b200 import _lib as L
    lib = L.load()
    cfg = L.MmifLossCfg(1.0, 0.01, 0.1, 1.0, 0, 0, 1, 1, 0, 0)
    assert lib.mmif_fusion_loss_fwd(None, None, None, 1, 32, 32, ctypes.byref(cfg), None, None, None, 0, None) == -1
    assert b'null' in lib.mmif_last_error()
    assert lib.mmif_fusion_loss_fwd(16, 16, 16, 1, 8, 32, ctypes.byref(cfg), 16, None, 16, 0, None) == -2   # H < 11
    assert lib.mmif_fusion_loss_fwd(18, 16, 16, 1, 32, 32, ctypes.byref(cfg), 16, None, 16, 0, None) == -5  # alignment
    bad = L.MmifLossCfg(1.0, 0.01, 0.1, 1.0, 7, 0, 1, 1, 0, 0)
    assert lib.mmif_fusion_loss_fwd(16, 16, 16, 1, 32, 32, ctypes.byref(bad), 16, None, 16, 0, None) == -3   # mode
    assert lib.mmif_loss_workspace_bytes(1, 8, 8) == 0
    assert lib.mmif_loss_workspace_bytes(64, 3072, 4096) > 0
    assert lib.mmif_loss_out_doubles(3) == (4 + 18) + (4 + 18 + 1) // 2          # the double block + its float32 mirror
    g = ctypes.c_float(1.0)
    assert lib.mmif_fusion_loss_bwd3(16, 16, 16, 1, 32, 32, ctypes.byref(cfg), None, None, None, None, 16, 16, 0, None) == -1   # no upstream
    assert lib.mmif_fusion_loss_bwd3(16, 16, 16, 1, 32, 32, ctypes.byref(cfg), 18, None, None, None, 16, 16, 0, None) == -5    # alignment
    assert lib.mmif_fusion_loss_bwd3(16, 16, 16, 1, 32, 32, ctypes.byref(cfg), 16, None, None, None, None, 16, 0, None) == -1   # null dF
    counts = (ctypes.c_ulonglong * 16)()
    assert lib.mmif_launch_counts(counts, 16) == 0 and lib.mmif_launch_counts(None, 16) == -1
    assert L.launch_counts()['loss_single_pass'] == 0          # nothing has been launched on this GPU-less host
    assert lib.mmif_metric_workspace_bytes(21, 480, 640) > 21 * (3 * 256 + 2 * 65536) * 4
    assert lib.mmif_set_gaussian_taps(18, 1.5, (ctypes.c_float * 18)()) == -2
    # SSIM(win_size) entries: window set, shape against the window, null pointers, workspace
    assert lib.mmif_ssim_fwd_win(None, None, None, 1, 32, 32, 7, 1.0, None, None, 0, None) == -1
    assert lib.mmif_ssim_fwd_win(16, 16, 16, 1, 32, 32, 8, 1.0, 16, 16, 0, None) == -3            # even window
    assert b'window' in lib.mmif_last_error()
    assert lib.mmif_ssim_fwd_win(16, 16, 16, 1, 4, 32, 5, 1.0, 16, 16, 0, None) == -2             # H < window
    assert lib.mmif_ssim_fwd_win(16, 16, 16, 1, 32, 32, 5, 1.0, 16, 16, 0, None) == -4            # workspace too small
    assert lib.mmif_ssim_bwd_ex_win(16, 16, 16, 1, 32, 32, 13, 1.0, 16, None, 0, 1.0, 16, 16, 0, None) == -3
    assert lib.mmif_ssim_bwd_ex_win(16, 16, 16, 1, 32, 32, 7, 1.0, None, None, 0, 1.0, 16, 16, 0, None) == -1


def test_dropin_surface_and_error_behaviour():
    import mmif_b200  # noqa: F401
    from mmif_b200.core import loss as ML, metric as MM
    assert ML.__all__ == ['SSIM', 'MS_SSIM', 'MSW_SSIM', 'SSIMLoss', 'PixelLoss', 'GradLoss', 'TVLoss', 'NormLoss']
    assert MM.__all__ == ['calc_mean', 'calc_std', 'calc_ag', 'calc_sf', 'calc_mse', 'calc_psnr', 'calc_cc', 'calc_scd',
                          'calc_entropy', 'calc_cross_ent', 'calc_mul_info', 'calc_Qabf', 'calc_Nabf', 'calc_Labf',
                          'calc_ssim', 'calc_msssim', 'calc_viff']
    assert set(ML.SSIM().state_dict()) == {'window'}
    assert set(ML.MS_SSIM().state_dict()) == {'window', 'weights'}
    assert set(ML.GradLoss().state_dict()) == {'x_sobel', 'y_sobel'}
    from oracle import fusion_loss as OL
    assert torch.equal(ML.SSIM().window, OL.window2d(11, 1.5))
    x = torch.rand(1, 1, 16, 16)
    with pytest.raises(ValueError):
        ML.SSIMLoss('nope')(x, x, x)
    with pytest.raises(ValueError):
        ML.NormLoss('l3')(x)
    with pytest.raises(Exception) as ei:          # no CPU compute path behind the drop-ins
        ML.SSIMLoss('ssim')(x, x, x)
    assert 'CUDA' in str(ei.value)
    # calc_psnr is scalar arithmetic on a tensor, identical to the reference
    assert abs(MM.calc_psnr(torch.tensor(0.01)).item() - 20.0) < 1e-5


def test_core_shim_resolves_new_loss_and_metric(tmp_path):
    """`from core.loss import ...` (train.py:31) hits the B200 package, other core.* modules fall
    through to whatever `core` directory follows on sys.path (the reference checkout)."""
    import subprocess
    import sys
    other = tmp_path / 'refrepo' / 'core'
    other.mkdir(parents=True)
    (other / 'model.py').write_text('MARK = "reference model module"\n')
    code = ("import sys; sys.path[:0] = [%r, %r]; import core.loss, core.metric, core.model; "
            "print(core.loss.SSIMLoss.__module__, core.model.MARK)") % (
        os.path.join(ROOT, 'dropin'), str(tmp_path / 'refrepo'))
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert 'reference model module' in out.stdout and 'core.loss' in out.stdout


def test_aggregate_quirk_matches_oracle():
    import mmif_b200  # noqa: F401
    from mmif_b200 import dist_utils as DU
    from oracle import fusion_metric as OM
    rng = np.random.default_rng(0)
    table = rng.random((5, 16))
    rows = [dict(zip(OM.METRIC_NAMES, r)) for r in table]
    assert DU.aggregate_columns(table) == OM.aggregate_columns(rows)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import sys
    sys.path.insert(0, ROOT)
    import mmif_b200  # noqa: F401
    from mmif_b200 import dist_utils as DU
    # (1) loss scalars: one packed all-reduce == four reduce_value calls
    vals = [torch.tensor(float(rank + 1) * v) for v in (1.0, 0.5, 0.25, 0.125)]
    packed = [t.item() for t in DU.reduce_loss_scalars(*vals, world)]
    ref = []
    for v in vals:
        t = v.clone()
        dist.all_reduce(t)
        ref.append((t / world).item())
    # (2) sharded evaluation with an injected row function (pair index encoded in the image)
    def load_pair(i):
        shape = (1, 1, 12, 16) if i % 3 else (1, 1, 10, 20)        # two shape groups
        a = torch.full(shape, float(i))
        return a, a + 1, a + 2

    def rows(a, b, f):
        base = a[:, 0, 0, 0].double()
        return torch.stack([base * 16 + k for k in range(16)], dim=1)
    cols = DU.evaluate_sharded(load_pair, 7, rank, world, compute_rows=rows)
    q.put((rank, packed, ref, cols))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_loss_reduction_and_sharded_eval():
    import mmif_b200  # noqa: F401
    from mmif_b200 import dist_utils as DU
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, packed, ref, cols in got:
        assert packed == ref
        if rank == 0:
            single = DU.evaluate_sharded(
                lambda i: ((torch.full((1, 1, 12, 16) if i % 3 else (1, 1, 10, 20), float(i)),) * 3), 7, 0, 1,
                compute_rows=lambda a, b, f: torch.stack([a[:, 0, 0, 0].double() * 16 + k for k in range(16)], dim=1))
            assert cols == single
        else:
            assert cols is None


def test_fastcall_binding_loads_and_validates_without_gpu():
    """The CPython fast-call binding (csrc/fastcall.c) resolves the same library entries as the ctypes binding: the same
    return codes for the same invalid arguments, no compute call."""
    import mmif_b200  # noqa: F401
    from mmif_b200 import _lib as L
    fc = L.fastcall()
    assert fc is not None, '_fastcall.so was not built (make)'
    lib = L.load()
    cfg = L.MmifLossCfg()
    cfg.pixel_combine = cfg.grad_combine = L.COMBINE['max']
    cfg.pixel_norm = cfg.grad_norm = L.NORM['l1']
    import ctypes
    addr = ctypes.addressof(cfg)
    rc_fast = fc.loss_fwd(None, None, None, 1, 32, 32, addr, None, None, None, 0, None)
    rc_ct = lib.mmif_fusion_loss_fwd(None, None, None, 1, 32, 32, ctypes.byref(cfg), None, None, None, 0, None)
    assert rc_fast == rc_ct != 0
    rc_fast = fc.loss_bwd3(16, 16, 16, 1, 4, 4, addr, None, None, None, None, None, None, 0, None)      # shape too small
    rc_ct = lib.mmif_fusion_loss_bwd3(16, 16, 16, 1, 4, 4, ctypes.byref(cfg), None, None, None, None, None, None, 0, None)
    assert rc_fast == rc_ct != 0
    with pytest.raises(TypeError):
        fc.loss_fwd(1, 2, 3)


def test_loss_geometry_owns_every_row_and_column_exactly_once():
    """Host-side check of the row-segment search (no GPU): for a sweep of shapes, decoding every CTA index the way the kernels
    do (two-level tall / short segments; for the warp-specialised kernel also the fine class of the last strips) must
    cover each (gradient row, gradient column) of a sample exactly once, with the CTA count the workspace is sized for."""
    import ctypes
    import numpy as np
    import mmif_b200  # noqa: F401
    from mmif_b200 import _lib as L
    lib = L.load()
    rng = np.random.RandomState(5)
    shapes = [(1, 11, 11), (2, 96, 160), (8, 256, 256), (1, 1024, 1224), (3, 517, 1030), (8, 3072, 4096), (64, 3072, 4096),
              (5, 2000, 109), (1, 4000, 6000), (7, 333, 5000)]
    shapes += [(int(rng.randint(1, 33)), int(rng.randint(11, 3000)), int(rng.randint(11, 5000))) for _ in range(40)]
    for (B, H, W) in shapes:
        for kernel in (0, 1):
            out = (ctypes.c_int * 10)()
            L.check(lib.mmif_loss_geometry(B, H, W, kernel, out))
            nstrip, nseg, n_tall, T, s, F, fr, nsf, ctas, tg = list(out)
            assert nstrip == -(-W // tg) and 0 <= F < max(nstrip, 1) and (kernel == 1 or F == 0)
            assert ctas == (nstrip - F) * nseg + F * nsf
            rows = np.zeros(H, dtype=np.int32)
            for seg in range(nseg):                                   # coarse strips: tall segments first, then short ones
                i0 = seg * T if seg < n_tall else n_tall * T + (seg - n_tall) * s
                h = T if seg < n_tall else s
                assert i0 < H, (B, H, W, kernel, list(out))           # no empty CTA
                rows[i0:min(i0 + h, H)] += 1
            assert (rows == 1).all(), (B, H, W, kernel, list(out))
            if F:
                rows[:] = 0
                for seg in range(nsf):
                    assert seg * fr < H
                    rows[seg * fr:min(seg * fr + fr, H)] += 1
                assert (rows == 1).all(), (B, H, W, kernel, list(out))
            assert T % 8 == 0 and s % 8 == 0 and (F == 0 or fr % 8 == 0)
