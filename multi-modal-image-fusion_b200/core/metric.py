"""Drop-in mirror of the reference ``core/metric.py`` call surface on the B200 kernels.

Same 17 function names, argument meaning and return types as the reference
(metric.py:16-21): every function returns a 0-dim tensor (``calc_mul_info`` in float64, the rest
float32) on the device of its inputs.  ``eval.py`` hands CPU tensors to these functions
(eval.py:198-200 discards the ``.to(device)`` results); they are uploaded once per image (cached
by tensor identity) and ALL arithmetic runs on the GPU — there is no CPU compute path.

``eval_metrics_batch`` is the batched entry (N pairs per launch) the sharded evaluation and the
benchmark use; the per-function drop-ins share kernel results through a small per-pair memo so
the 20-odd calls of ``eval_metrics`` (eval.py:29-75) cost one launch per kernel family.
"""
import ctypes
import math
import threading
import weakref

import torch

from .. import _lib as L

__all__ = [
    'calc_mean', 'calc_std', 'calc_ag', 'calc_sf', 'calc_mse', 'calc_psnr',
    'calc_cc', 'calc_scd', 'calc_entropy', 'calc_cross_ent', 'calc_mul_info',
    'calc_Qabf', 'calc_Nabf', 'calc_Labf', 'calc_ssim', 'calc_msssim',
    'calc_viff'
]

METRIC_NAMES = ('sd', 'ag', 'sf', 'mse', 'psnr', 'cc', 'scd', 'en', 'ce', 'mi',
                'qabf', 'nabf', 'labf', 'ssim', 'msssim', 'viff')   # eval.py:52-68


def _default_device():
    if not torch.cuda.is_available():
        raise L.MmifError('no CUDA device: the metric suite has no CPU compute path')
    return torch.device('cuda', torch.cuda.current_device())


class _UploadCache:
    """CPU tensor -> device copy, keyed on identity + version (eval.py passes the same three CPU
    tensors to ~20 metric calls)."""

    def __init__(self, cap=8):
        self.cap, self.items = cap, []

    def get(self, t):
        if t.is_cuda:
            return t
        for ref, ver, dev in self.items:
            if ref() is t and ver == t._version:
                return dev
        dev = _staged_upload(t, _default_device())
        self.items.append((weakref.ref(t), t._version, dev))
        if len(self.items) > self.cap:
            self.items.pop(0)
        return dev


class _PinnedStage:
    """Pageable host tensor -> device through a small ring of pinned staging buffers.  `tensor.to(device)` from pageable
    memory costs ~0.4 ms per call on this platform whatever the size (measured: 0.41 ms for 1.2 MB, 0.50 ms for 5 MB;
    tools/upload_test.py), the host copy into pinned memory (torch parallelises it) + an asynchronous DMA 0.15 ms: the three
    images eval.py hands over per pair (eval.py:189-200 keeps them on the CPU) upload in 0.45 instead of 1.2 - 1.5 ms."""

    def __init__(self, slots=3):
        self.bufs, self.events, self.k = [None] * slots, [None] * slots, 0
        self.lock = threading.Lock()

    def upload(self, t, device):
        nbytes = t.numel() * t.element_size()
        # (with one host thread — torchrun sets OMP_NUM_THREADS=1 — the host copy is no faster than the driver's own staging)
        if (t.is_pinned() or nbytes < (64 << 10) or nbytes > (256 << 20) or t.is_sparse or t.layout != torch.strided
                or torch.get_num_threads() < 4):
            return t.to(device, non_blocking=False)
        with self.lock:
            k = self.k
            self.k = (k + 1) % len(self.bufs)
            if self.events[k] is not None:
                self.events[k].synchronize()              # the DMA that last read this staging buffer has finished
            if self.bufs[k] is None or self.bufs[k].numel() < nbytes:
                self.bufs[k] = torch.empty(max(nbytes, 8 << 20), dtype=torch.uint8, pin_memory=True)
            stage = self.bufs[k][:nbytes].view(t.dtype).view(t.shape)
            stage.copy_(t)
            with torch.cuda.device(device):
                dev = stage.to(device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
            self.events[k] = ev
        return dev


_stage = _PinnedStage()


def _staged_upload(t, device):
    return _stage.upload(t, device)


_uploads = _UploadCache()


def _prep(*imgs):
    """-> list of contiguous float32 (N,H,W) device views + (N,H,W) + the device results go back to."""
    home = imgs[0].device
    out, shape = [], None
    for i, t in enumerate(imgs):
        d = _uploads.get(t)
        d, n, h, w = L.as_f32_3d(d, f'img{i}')
        if shape is None:
            shape = (n, h, w)
        elif shape != (n, h, w):
            raise L.MmifError(f'shape mismatch: {shape} vs {(n, h, w)}')
        out.append(d)
    L.ensure_device(out[0].device)
    return out, shape, home


class _PairMemo:
    """Results of one kernel family for one (a, b, f) triple, keyed on storage identity/version.  A result whose caller
    lives on the CPU (eval.py) is copied to the host ONCE per family (`host`), so the ~25 calc_* calls of one pair cost one
    device->host transfer + sync per kernel family instead of one per returned scalar."""

    def __init__(self, cap=12):
        self.cap, self.items = cap, []

    @staticmethod
    def _key(kind, tensors, extra):
        return (kind, extra) + tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in tensors)

    def get(self, kind, tensors, extra, compute):
        key = self._key(kind, tensors, extra)
        for ent in self.items:
            if ent[0] == key and all(r() is t for r, t in zip(ent[1], tensors)):
                return ent
        ent = [key, tuple(weakref.ref(t) for t in tensors), compute(), None]
        self.items.append(ent)
        if len(self.items) > self.cap:
            self.items.pop(0)
        return ent


_memo = _PairMemo()


def _combine_batch(kind, val, part, extra):
    """Per-pair rows (n, K) -> ONE row (1, K) in the same layout holding what the reference's function returns for a batch:
    every reference metric reduces over ALL dimensions, i.e. over the whole (N,1,H,W) batch as one population
    (metric.py:25-491) — global means / variances / correlations, one histogram, ratios of batch-wide sums.  float64 torch
    ops on the kernels' per-pair results (a few hundred scalars; the pixels were reduced by the kernels)."""
    rows = val if part is None else val[part]
    n = rows.shape[0] if rows is not None else 0
    if kind == 'stats':
        r = rows
        mf, ma, mb = r[:, 0].mean(), r[:, 9].mean(), r[:, 10].mean()
        e2 = lambda sd, m: (sd * sd + m * m).mean()                       # E[x^2] over the batch
        vf, va, vb = e2(r[:, 1], r[:, 0]) - mf * mf, e2(r[:, 11], r[:, 9]) - ma * ma, e2(r[:, 12], r[:, 10]) - mb * mb
        cov = lambda cc, s1, s2, m1, m2, g1, g2: (cc * s1 * s2 + m1 * m2).mean() - g1 * g2
        caf = cov(r[:, 6], r[:, 11], r[:, 1], r[:, 9], r[:, 0], ma, mf)
        cbf = cov(r[:, 7], r[:, 12], r[:, 1], r[:, 10], r[:, 0], mb, mf)
        cab = cov(r[:, 13], r[:, 11], r[:, 12], r[:, 9], r[:, 10], ma, mb)
        out = torch.zeros(1, L.ST_COUNT, dtype=torch.float64, device=r.device)
        out[0, 0], out[0, 1], out[0, 2] = mf, vf.clamp(min=0).sqrt(), r[:, 2].mean()
        out[0, 3] = (r[:, 3] * r[:, 3]).mean().sqrt()                      # sf^2 = mean dv^2 + mean dh^2 is additive over the batch
        out[0, 4], out[0, 5] = r[:, 4].mean(), r[:, 5].mean()
        out[0, 6], out[0, 7] = caf / (va * vf).sqrt(), cbf / (vb * vf).sqrt()
        v_fa, v_fb = vf + va - 2.0 * caf, vf + vb - 2.0 * cbf             # scd = cc(f-a, b) + cc(f-b, a), metric.py:95-99
        out[0, 8] = (cbf - cab) / (v_fa * vb).sqrt() + (caf - cab) / (v_fb * va).sqrt()
        out[0, 9], out[0, 10], out[0, 11], out[0, 12] = ma, mb, va.clamp(min=0).sqrt(), vb.clamp(min=0).sqrt()
        out[0, 13] = cab / (va * vb).sqrt()
        return out
    if kind == 'hist':                                                    # ONE histogram over the batch (metric.py:103-188)
        c = val[0].to(torch.int64).sum(dim=0)
        total = float(extra)                                              # pixels in the batch
        ha, hb, hf = (c[k * 256:(k + 1) * 256].double() for k in range(3))
        jaf, jbf = c[768:768 + 65536].double(), c[768 + 65536:].double()

        def ent(h):
            p = h[h > 0] / total
            return -(p * torch.log2(p)).sum()

        def cross(h1, h2):
            m = (h1 > 0) & (h2 > 0)
            return (h1[m] / total * torch.log2(h1[m] / h2[m])).sum()

        ea, eb, ef, jea, jeb = ent(ha), ent(hb), ent(hf), ent(jaf), ent(jbf)
        out = torch.zeros(1, L.EN_COUNT, dtype=torch.float64, device=c.device)
        out[0, 0], out[0, 1], out[0, 2], out[0, 3], out[0, 4] = ea, eb, ef, jea, jeb
        out[0, 5], out[0, 6] = cross(ha, hf), cross(hb, hf)
        out[0, 7], out[0, 8] = ea + ef - jea, eb + ef - jeb
        out[0, 9], out[0, 10] = 2.0 * out[0, 7] / (ea + ef), 2.0 * out[0, 8] / (eb + ef)
        return out
    if kind == 'qabf':                                                    # rows = mmif_qabf_raw: ratios of batch-wide sums
        t = rows[:, 4:9].sum(dim=0)
        return torch.stack([t[0] / t[1], t[2] / t[1], t[3] / t[1], t[4] / t[1]]).view(1, 4)
    if kind == 'ssim':                                                    # global mean of the maps (metric.py:357-364)
        return rows.mean(dim=0, keepdim=True)
    if kind == 'msssim':                                                  # per-level global means, then the product (metric.py:368-402)
        lv = rows[:, 2:22].view(n, 5, 4).mean(dim=0)                      # (5, [ssim_af, cs_af, ssim_bf, cs_bf])
        wts = torch.tensor([0.0448, 0.2856, 0.3001, 0.2363, 0.1333], dtype=torch.float64, device=rows.device)
        out = torch.zeros(1, L.MSSSIM_DOUBLES, dtype=torch.float64, device=rows.device)
        for k, (cs_col, ss_col) in enumerate(((1, 0), (3, 2))):
            v = torch.cat([lv[:4, cs_col], lv[4:, ss_col]]).clamp(min=1e-7)
            out[0, k] = torch.prod(v ** wts)
        out[0, 2:] = lv.reshape(-1)
        return out
    if kind == 'viff':                                                    # sums over the batch per scale (metric.py:461-491)
        sc = rows[:, 2:26].view(n, 4, 6).sum(dim=0)                       # (4, [num1, den1, num2, den2, num_sel, den_sel])
        p = torch.tensor([1.0, 0.0, 0.15, 1.0], dtype=torch.float64, device=rows.device) / 2.15
        out = torch.zeros(1, L.VIFF_DOUBLES, dtype=torch.float64, device=rows.device)
        out[0, 0] = (p * sc[:, 4] / sc[:, 5]).sum()
        out[0, 1] = sc[:, 0].sum() / sc[:, 1].sum() + sc[:, 2].sum() / sc[:, 3].sum()
        out[0, 2:] = sc.reshape(-1)
        return out
    raise L.MmifError(f'no batch rule for {kind}')


def _rows(ent, home, part=None):
    """The float64 row(s) of a memo entry where the caller lives: the device tensor, or its single host copy.  A batch
    (N > 1) is first combined into the ONE row the reference's function returns for it (`_combine_batch`)."""
    val = ent[2] if part is None else ent[2][part]
    n = (ent[2][1] if ent[0][0] == 'hist' else val).shape[0]
    if ent[3] is None:
        ent[3] = {}
    if n > 1:
        if ('g', part) not in ent[3]:
            ent[3][('g', part)] = _combine_batch(ent[0][0], ent[2], part, math.prod(ent[0][-1][2]))
        val = ent[3][('g', part)]
    if home.type != 'cpu':
        return val if val.device == home else val.to(home)
    if ('h', part) not in ent[3]:
        ent[3][('h', part)] = val.cpu()
    return ent[3][('h', part)]


class _Triples:
    """eval.py:29-75 calls the one- and two-image functions (calc_std(f), calc_mse(a, f), calc_mse(b, f), calc_cc, ...) one
    after another on the SAME three tensors.  The kernels are three-image kernels, so once the triple (a, b, f) is known
    — from any three-image call, or from two two-image calls (x, f), (x', f) on one fused image — every later call on its
    members is served by ONE launch per kernel family on (a, b, f) instead of one launch per distinct argument list."""

    def __init__(self, cap=4):
        self.cap, self.triples, self.pending = cap, [], []       # entries: tuples of (weakref, version)

    @staticmethod
    def _same(entry, t):
        return entry[0]() is t and entry[1] == t._version

    @staticmethod
    def _mk(t):
        return (weakref.ref(t), t._version)

    def add(self, a, b, f):
        for tr in self.triples:
            if self._same(tr[0], a) and self._same(tr[1], b) and self._same(tr[2], f):
                return
        self.triples.append((self._mk(a), self._mk(b), self._mk(f)))
        if len(self.triples) > self.cap:
            self.triples.pop(0)

    def _live(self, tr):
        t = tuple(e[0]() for e in tr)
        return t if all(x is not None and x._version == e[1] for x, e in zip(t, tr)) else None

    def of_fused(self, f):
        """-> (a, b, f) of the most recent triple whose fused image is `f`, or None."""
        for tr in reversed(self.triples):
            if self._same(tr[2], f):
                t = self._live(tr)
                if t is not None:
                    return t
        return None

    def of_pair(self, x, f):
        """(x, f) -> ((a, b, f), slot) with x == a (slot 0) or x == b (slot 1); learns the triple from the second distinct
        source seen with the same fused image; (None, 0) while only one source is known."""
        if x.shape != f.shape:
            return None, 0
        for tr in reversed(self.triples):
            if self._same(tr[2], f) and (self._same(tr[0], x) or self._same(tr[1], x)):
                t = self._live(tr)
                if t is not None:
                    return t, (0 if t[0] is x else 1)
        for px, pf in reversed(self.pending):
            if self._same(pf, f) and not self._same(px, x):
                other = px[0]()
                if other is not None and other._version == px[1] and other.shape == x.shape and other.device == x.device:
                    self.add(other, x, f)
                    return (other, x, f), 1
        self.pending.append((self._mk(x), self._mk(f)))
        if len(self.pending) > self.cap:
            self.pending.pop(0)
        return None, 0


_triples = _Triples()


def _ws(dev, n, h, w):
    lib = L.load()
    nbytes = lib.mmif_metric_workspace_bytes(n, h, w)
    if nbytes == 0:
        raise L.MmifError(f'unsupported shape {(n, h, w)}')
    return L.workspace(dev, nbytes, 'metric', (n, h, w))


def _call(fn_name, imgs, shape, out_doubles, *extra_args):
    lib = L.load()
    n, h, w = shape
    dev = imgs[0].device
    out = torch.empty(n * out_doubles, dtype=torch.float64, device=dev)
    ws = _ws(dev, n, h, w)
    L.call(dev, getattr(lib, fn_name), imgs[0].data_ptr(), imgs[1].data_ptr(), imgs[2].data_ptr(), n, h, w, *extra_args,
           out.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_int(dev))
    return out.view(n, out_doubles)


def _stats(a, b, f):
    def run():
        imgs, shape, _ = _prep(a, b, f)
        return _call('mmif_stats', imgs, shape, L.ST_COUNT)
    return _memo.get('stats', (a, b, f), None, run)


def _hist(a, b, f):
    """memo entry whose value is (counts (n, HIST_WORDS) int32, entropies (n, EN_COUNT) float64)."""
    def run():
        lib = L.load()
        imgs, (n, h, w), _ = _prep(a, b, f)
        dev = imgs[0].device
        counts = torch.empty(n * L.HIST_WORDS, dtype=torch.int32, device=dev)
        ent = torch.empty(n * L.EN_COUNT, dtype=torch.float64, device=dev)
        ws = _ws(dev, n, h, w)
        L.call(dev, lib.mmif_hist, imgs[0].data_ptr(), imgs[1].data_ptr(), imgs[2].data_ptr(), n, h, w, counts.data_ptr(),
               ent.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_int(dev))
        return counts.view(n, L.HIST_WORDS), ent.view(n, L.EN_COUNT)
    return _memo.get('hist', (a, b, f), None, run)


def hist_raw(a, b, f):
    """(counts (n, MMIF_HIST_WORDS) int32, entropies (n, MMIF_EN_COUNT) float64) device tensors of mmif_hist."""
    return _hist(a, b, f)[2]


def _qabf(a, b, f, Lexp):
    def run():
        imgs, shape, _ = _prep(a, b, f)
        return _call('mmif_qabf_raw', imgs, shape, 9, ctypes.c_float(Lexp))
    _triples.add(a, b, f)
    return _memo.get('qabf', (a, b, f), float(Lexp), run)


def _one(f):
    """Arguments for a one-image call on `f`: the known triple that has it as its fused image, else (f, f, f)."""
    tr = _triples.of_fused(f)
    return tr if tr is not None else (f, f, f)


def _two(x, f):
    """Arguments + slot for a two-image call (x, f): the known triple (slot 0: x is its first source, 1: its second),
    else (x, x, f) with slot 0."""
    tr, slot = _triples.of_pair(x, f)
    return (tr, slot) if tr is not None else ((x, x, f), 0)


def _pad_dev(t, pad):
    """reflect-pad (N,H,W) device images by `pad` (use_padding=True, metric.py:305-311)."""
    lib = L.load()
    n, h, w = t.shape
    if pad >= h or pad >= w:
        raise L.MmifError(f'reflect padding {pad} needs H, W > {pad}, got {(h, w)} (torch raises here too)')
    out = torch.empty(n, h + 2 * pad, w + 2 * pad, dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        L.check(lib.mmif_reflect_pad(t.data_ptr(), n, h, w, pad, out.data_ptr(), L.stream_ptr(t.device)))
    return out


def _halve_dev(t):
    lib = L.load()
    n, h, w = t.shape
    out = torch.empty(n, (h + 1) // 2, (w + 1) // 2, dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        L.check(lib.mmif_halve(t.data_ptr(), n, h, w, out.data_ptr(), L.stream_ptr(t.device)))
    return out


def _ssim_level(imgs, shape, win_size, data_range, use_padding):
    """(n,4) [ssim_af, cs_af, ssim_bf, cs_bf] of one level; use_padding pads by min(win,h,w)//2 first."""
    n, h, w = shape
    if use_padding:
        p = min(int(win_size), h, w) // 2
        if p:
            imgs = [_pad_dev(t.view(n, h, w), p) for t in imgs]
            shape = (n, h + 2 * p, w + 2 * p)
    return _call('mmif_ssim', imgs, shape, 4, int(win_size) if not use_padding else min(int(win_size), h, w),
                 ctypes.c_float(data_range), 0)


def _ssim2(a, b, f, win_size, data_range, use_padding):
    def run():
        imgs, shape, _ = _prep(a, b, f)
        return _ssim_level(imgs, shape, win_size, data_range, use_padding)
    return _memo.get('ssim', (a, b, f), (int(win_size), float(data_range), bool(use_padding)), run)


def _msssim2(a, b, f, win_size, data_range, use_padding):
    def run():
        imgs, shape, _ = _prep(a, b, f)
        if not use_padding:
            return _call('mmif_msssim', imgs, shape, L.MSSSIM_DOUBLES, int(win_size), ctypes.c_float(data_range))
        # use_padding=True pads every level before its blur while the pyramid pools the unpadded level
        # (metric.py:378-400): composed level by level from the same device kernels
        n, h, w = shape
        wts = torch.tensor([0.0448, 0.2856, 0.3001, 0.2363, 0.1333], dtype=torch.float64, device=imgs[0].device)
        cur = [t.view(n, h, w) for t in imgs]
        vals, levels = [], []
        for lvl in range(5):
            hh, ww = cur[0].shape[-2:]
            r = _ssim_level(cur, (n, hh, ww), win_size, data_range, True)     # window = min(win, h, w) per level (metric.py:323-325)
            vals.append(torch.stack([r[:, 1], r[:, 3]], dim=1) if lvl < 4 else torch.stack([r[:, 0], r[:, 2]], dim=1))
            levels.append(r)
            if lvl < 4:
                cur = [_halve_dev(t) for t in cur]
        v = torch.stack(vals, dim=0).clamp(min=1e-7)       # (5, n, 2)
        ms = torch.prod(v ** wts.view(5, 1, 1), dim=0)
        out = torch.zeros(n, L.MSSSIM_DOUBLES, dtype=torch.float64, device=ms.device)
        out[:, 0], out[:, 1] = ms[:, 0], ms[:, 1]
        out[:, 2:] = torch.stack(levels, dim=1).reshape(n, 20)          # per level ssim_af, cs_af, ssim_bf, cs_bf (mmif_msssim layout)
        return out
    return _memo.get('msssim', (a, b, f), (int(win_size), float(data_range), bool(use_padding)), run)


def _viff3(a, b, f):
    _triples.add(a, b, f)

    def run():
        imgs, shape, _ = _prep(a, b, f)
        return _call('mmif_viff', imgs, shape, L.VIFF_DOUBLES)
    return _memo.get('viff', (a, b, f), None, run)


def _scalar(rows, idx, dtype=torch.float32):
    """0-dim tensor rows[0, idx] in the reference's dtype, on the device `rows` lives on (the caller's)."""
    return rows[0, idx].to(dtype)


def _single(t, what):
    """Every reference metric reduces over the whole batch; N > 1 is served by the batched kernels + `_combine_batch`."""
    if t.dim() not in (2, 3, 4):
        raise L.MmifError(f'{what}: (N,1,H,W), (N,H,W) or (H,W) expected, got {tuple(t.shape)}')


def _stat1(img, idx_f, idx_a, idx_b):
    """A one-image statistic of `img`: column idx_f of the triple that has it as fused image, else of (img, img, img)."""
    tr = _one(img)
    return _scalar(_rows(_stats(*tr), img.device), idx_f)


# 1. mean
def calc_mean(img):
    _single(img, 'calc_mean')
    return _stat1(img, 0, 9, 10)


# 2. sd
def calc_std(img):
    _single(img, 'calc_std')
    return _stat1(img, 1, 11, 12)


# 3. ag
def calc_ag(img):
    _single(img, 'calc_ag')
    return _stat1(img, 2, None, None)


# 4. sf
def calc_sf(img):
    _single(img, 'calc_sf')
    return _stat1(img, 3, None, None)


# 5. mse  (kernel slots: a vs f = column 4, b vs f = column 5)
def calc_mse(img1, img2):
    _single(img1, 'calc_mse')
    tr, slot = _two(img1, img2)
    return _scalar(_rows(_stats(*tr), img1.device), 4 + slot)


# 6. psnr — pure scalar arithmetic on the mse tensor, as in the reference (metric.py:72-76)
def calc_psnr(mse, L=1.0, root=False):
    if root:
        return 20.0 * torch.log10(L / mse**0.5)
    return 10.0 * torch.log10(L**2 / mse)


# 7. cc
def calc_cc(img1, img2):
    _single(img1, 'calc_cc')
    tr, slot = _two(img1, img2)
    return _scalar(_rows(_stats(*tr), img1.device), 6 + slot)


# 8. scd
def calc_scd(img1, img2, imgf):
    _single(img1, 'calc_scd')
    _triples.add(img1, img2, imgf)
    return _scalar(_rows(_stats(img1, img2, imgf), img1.device), 8)


# 9. en
def calc_entropy(img):
    _single(img, 'calc_entropy')
    tr = _one(img)
    return _scalar(_rows(_hist(*tr), img.device, 1), 2)


# 11. ce
def calc_cross_ent(img1, img2):
    _single(img1, 'calc_cross_ent')
    tr, slot = _two(img1, img2)
    return _scalar(_rows(_hist(*tr), img1.device, 1), 5 + slot)


# 12. mi — float64 like the reference (metric.py:179-188)
def calc_mul_info(img1, img2, normalized=False):
    _single(img1, 'calc_mul_info')
    tr, slot = _two(img1, img2)
    return _scalar(_rows(_hist(*tr), img1.device, 1), (9 if normalized else 7) + slot, torch.float64)


# 13. Qabf
def calc_Qabf(img1, img2, imgf, L=1.5, full=False):
    _single(img1, 'calc_Qabf')
    q = _rows(_qabf(img1, img2, imgf, L), img1.device)
    if full:
        return tuple(_scalar(q, i) for i in range(3))
    return _scalar(q, 0)


# 14. Nabf
def calc_Nabf(img1, img2, imgf, L=1.5, modified=True):
    _single(img1, 'calc_Nabf')
    return _scalar(_rows(_qabf(img1, img2, imgf, L), img1.device), 1 if modified else 3)


# 15. Labf
def calc_Labf(img1, img2, imgf, L=1.5):
    _single(img1, 'calc_Labf')
    return _scalar(_rows(_qabf(img1, img2, imgf, L), img1.device), 2)


# 16. ssim
def calc_ssim(img1, img2, win_size=11, data_range=255.0, use_padding=False, size_average=True, full=False):
    _single(img1, 'calc_ssim')
    if not size_average:
        return _ssim_map(img1, img2, win_size, data_range, use_padding, full)
    tr, slot = _two(img1, img2)
    r = _rows(_ssim2(*tr, win_size, data_range, use_padding), img1.device)
    if full:
        return _scalar(r, 2 * slot), _scalar(r, 2 * slot + 1)
    return _scalar(r, 2 * slot)


def _ssim_map(img1, img2, win_size, data_range, use_padding, full):
    """size_average=False (metric.py:357-364): the (1,1,H',W') SSIM map (and CS map with full=True)."""
    lib = L.load()
    imgs, (n, h, w), home = _prep(img1, img2)
    if min(int(win_size), h, w) != 11:
        raise NotImplementedError('SSIM maps are built for the 11-tap window (images of at least 11 x 11)')
    if use_padding:
        imgs = [_pad_dev(t.view(n, h, w), 5) for t in imgs]
        h, w = h + 10, w + 10
    dev = imgs[0].device
    nws = lib.mmif_loss_workspace_bytes(n, h, w)
    ws = L.workspace(dev, nws, 'loss', (n, h, w))
    maps = [torch.empty(n, 1, h - 10, w - 10, dtype=torch.float32, device=dev) for _ in range(2)]
    with torch.cuda.device(dev):
        L.check(lib.mmif_ssim_maps(imgs[0].data_ptr(), imgs[0].data_ptr(), imgs[1].data_ptr(), n, h, w, float(data_range),
                                   maps[0].data_ptr(), maps[1].data_ptr(), None, None, None, None, ws.data_ptr(), ws.numel(),
                                   L.stream_ptr(dev)))
    maps = [m.to(home) if m.device != home else m for m in maps]
    return (maps[0], maps[1]) if full else maps[0]


# 17. msssim
def calc_msssim(img1, img2, win_size=11, data_range=255.0, use_padding=False):
    _single(img1, 'calc_msssim')
    tr, slot = _two(img1, img2)
    return _scalar(_rows(_msssim2(*tr, win_size, data_range, use_padding), img1.device), slot)


# 18. viff
def calc_viff(img1, img2, imgf, simple=True):
    _single(img1, 'calc_viff')
    return _scalar(_rows(_viff3(img1, img2, imgf), img1.device), 1 if simple else 0)


# ---- batched entries (not in the reference; used by the sharded evaluation and the benchmark) ----
def eval_metrics_batch(img1, img2, imgf):
    """(N,1,H,W) x3 on the GPU -> (N,16) float64 tensor, columns in eval.py:52-68 order."""
    for t in (img1, img2, imgf):
        L.require_cuda(t, 'image')
    imgs, shape, _ = _prep(img1, img2, imgf)
    return _call('mmif_eval_suite', imgs, shape, L.EVAL_METRICS)


def test_post_step(img1, img2, imgf, data_range=1.0, want_image=True):
    """test.py:49-73 after the model call, in one launch and one pass over imgf:
    ``(ssim1 + ssim2) * 0.5`` with ``ssim_k = calc_ssim(img_k, imgf, data_range=1.0)`` per sample -> (B,) float32
    tensor, and ``save_result(imgf[b])`` = ``denorm`` (uint8(clip(0,1)*255), data/transform.py:32-35) -> (B,H,W) uint8
    device tensor (``.cpu().numpy()[b][..., None]`` is what cv2.imwrite gets in test.py:72-73)."""
    lib = L.load()
    for t in (img1, img2, imgf):
        L.require_cuda(t, 'image')
    imgs, (n, h, w), _ = _prep(img1, img2, imgf)
    dev = imgs[0].device
    nws = lib.mmif_loss_workspace_bytes(n, h, w)
    if nws == 0:
        raise L.MmifError(f'unsupported shape {(n, h, w)}: H and W must be >= 11')
    ws = L.workspace(dev, nws, 'loss', (n, h, w))
    out = torch.empty(lib.mmif_loss_out_doubles(n), dtype=torch.float64, device=dev)
    img8 = torch.empty(n, h, w, dtype=torch.uint8, device=dev) if want_image else None
    with torch.cuda.device(dev):
        L.check(lib.mmif_test_post(imgs[0].data_ptr(), imgs[1].data_ptr(), imgs[2].data_ptr(), n, h, w, float(data_range),
                                   out.data_ptr(), img8.data_ptr() if want_image else None, ws.data_ptr(), ws.numel(),
                                   L.stream_ptr(dev)))
    ps = out[L.LOSS_HEAD:L.LOSS_HEAD + n * L.LOSS_PER_SAMPLE].view(n, L.LOSS_PER_SAMPLE)
    return ((ps[:, 0] + ps[:, 3]) * 0.5).to(torch.float32), img8


def eval_metrics_batch_u8(img1, img2, imgf):
    """uint8 ingest: (N,1,H,W) / (N,H,W) uint8 tensors — CUDA, or pinned / pageable HOST tensors (one H2D copy of
    1 byte per pixel per image) -> (N,16) float64 rows on the GPU.  Same results as widening on the host
    (eval.py:182-194) followed by eval_metrics_batch: the device widening is exact."""
    lib = L.load()
    ts = []
    for t in (img1, img2, imgf):
        if t.dtype != torch.uint8:
            raise L.MmifError(f'uint8 expected, got {t.dtype}')
        if t.dim() == 4:
            if t.shape[1] != 1:
                raise L.MmifError(f'single-channel (N,1,H,W) expected, got {tuple(t.shape)}')
            t = t[:, 0]
        elif t.dim() == 2:
            t = t.unsqueeze(0)
        ts.append(t.contiguous())
    n, h, w = ts[0].shape
    if any(tuple(t.shape) != (n, h, w) for t in ts):
        raise L.MmifError('shape mismatch between the three images')
    dev = ts[0].device if ts[0].is_cuda else _default_device()
    L.ensure_device(dev)
    ts = [t.to(dev, non_blocking=True) for t in ts]
    scratch = torch.empty(3 * n * h * w, dtype=torch.float32, device=dev)
    out = torch.empty(n * L.EVAL_METRICS, dtype=torch.float64, device=dev)
    ws = _ws(dev, n, h, w)
    with torch.cuda.device(dev):
        L.check(lib.mmif_eval_suite_u8(ts[0].data_ptr(), ts[1].data_ptr(), ts[2].data_ptr(), n, h, w, out.data_ptr(),
                                       scratch.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr(dev)))
    return out.view(n, L.EVAL_METRICS)


_copy_streams = {}


def eval_metrics_batch_host(img1, img2, imgf, chunks=None):
    """HOST images (float32 as eval.py:189-194 builds them, or uint8 as cv2 decodes them, eval.py:182-187; pinned memory
    for real overlap) -> (N,16) float64 rows on the GPU, with the upload pipelined against the suite: the N pairs go in
    ``chunks`` pieces through two device staging buffers, piece k+1 copies on a side stream while piece k computes
    (default: about 24 MB of upload per piece, 2..8 pieces — measured best for 21 x 640x480 and 32 x 1224x1024 pairs).
    Rows equal ``eval_metrics_batch`` up to the fp32 summation order of a different batch split (~2e-6 relative, the
    histogram metrics exactly).  Asynchronous like the other entries: the caller's ``.cpu()`` / ``.item()`` syncs."""
    lib = L.load()
    ts = []
    for t in (img1, img2, imgf):
        if t.is_cuda:
            raise L.MmifError('eval_metrics_batch_host takes host tensors (use eval_metrics_batch / _u8 for device tensors)')
        if t.dtype not in (torch.uint8, torch.float32) or t.dtype != img1.dtype:
            raise L.MmifError(f'uint8 or float32 expected (all three alike), got {t.dtype}')
        if t.dim() == 4:
            if t.shape[1] != 1:
                raise L.MmifError(f'single-channel (N,1,H,W) expected, got {tuple(t.shape)}')
            t = t[:, 0]
        elif t.dim() == 2:
            t = t.unsqueeze(0)
        ts.append(t.contiguous())
    n, h, w = ts[0].shape
    if any(tuple(t.shape) != (n, h, w) for t in ts):
        raise L.MmifError('shape mismatch between the three images')
    is_u8 = ts[0].dtype == torch.uint8
    dev = _default_device()
    L.ensure_device(dev)
    if chunks is None:
        total = 3 * n * h * w * ts[0].element_size()
        chunks = max(2, min(8, -(-total // (24 << 20))))
    nchunk = max(1, min(int(chunks), n))
    base, extra = divmod(n, nchunk)
    sizes = [base + (1 if c < extra else 0) for c in range(nchunk)]
    comp = torch.cuda.current_stream(dev)
    copy = _copy_streams.get(dev.index)
    if copy is None:
        copy = _copy_streams[dev.index] = torch.cuda.Stream(device=dev)
    out = torch.empty(n * L.EVAL_METRICS, dtype=torch.float64, device=dev)
    stage = [[torch.empty((sizes[0], h, w), dtype=ts[0].dtype, device=dev) for _ in range(3)] for _ in range(min(2, nchunk))]
    scratch = torch.empty(3 * sizes[0] * h * w, dtype=torch.float32, device=dev) if is_u8 else None
    ready = [torch.cuda.Event() for _ in stage]
    free = [torch.cuda.Event() for _ in stage]
    copy.wait_stream(comp)              # the staging blocks may be recycled memory of earlier work on this stream
    lo = 0
    with torch.cuda.device(dev):
        for c, nc in enumerate(sizes):
            s = c & 1
            with torch.cuda.stream(copy):
                if c >= 2:
                    copy.wait_event(free[s])
                for k in range(3):
                    stage[s][k][:nc].copy_(ts[k][lo:lo + nc], non_blocking=True)
                ready[s].record(copy)
            comp.wait_event(ready[s])
            ws = _ws(dev, nc, h, w)
            optr = out.data_ptr() + lo * L.EVAL_METRICS * 8
            p = [t.data_ptr() for t in stage[s]]
            if is_u8:
                L.check(lib.mmif_eval_suite_u8(p[0], p[1], p[2], nc, h, w, optr, scratch.data_ptr(), ws.data_ptr(), ws.numel(),
                                               L.stream_ptr(dev)))
            else:
                L.check(lib.mmif_eval_suite(p[0], p[1], p[2], nc, h, w, optr, ws.data_ptr(), ws.numel(), L.stream_ptr(dev)))
            free[s].record(comp)
            lo += nc
    return out.view(n, L.EVAL_METRICS)


def eval_subset_batch(img1, img2, imgf, L_exp=1.5):
    """BASELINE configs[3] (polarization evaluation): MS-SSIM + VIFF + Qabf/Nabf/Labf only, for N pairs on the GPU ->
    (N,5) float64 [msssim, viff, qabf, nabf, labf] with eval.py's conventions (msssim = mean of the two pairs,
    viff simple=False, calc_Qabf(L=1.5, full=True)); three kernel families, 51.6 algorithmic bytes per pixel."""
    for t in (img1, img2, imgf):
        L.require_cuda(t, 'image')
    imgs, shape, _ = _prep(img1, img2, imgf)
    ms = _call('mmif_msssim', imgs, shape, L.MSSSIM_DOUBLES, 11, ctypes.c_float(255.0))
    vf = _call('mmif_viff', imgs, shape, L.VIFF_DOUBLES)
    q = _call('mmif_qabf', imgs, shape, 4, ctypes.c_float(L_exp))
    return torch.stack([(ms[:, 0] + ms[:, 1]) * 0.5, vf[:, 0], q[:, 0], q[:, 1], q[:, 2]], dim=1)


def eval_metrics(img1, img2, imgf):
    """The dict eval.py:29-75 builds for one pair (python floats), through the fused suite entry."""
    imgs, shape, _ = _prep(img1, img2, imgf)
    if shape[0] != 1:
        raise L.MmifError('eval_metrics takes one pair; use eval_metrics_batch for N pairs')
    row = _call('mmif_eval_suite', imgs, shape, L.EVAL_METRICS)[0].tolist()
    return dict(zip(METRIC_NAMES, row))


def histograms(img1, img2, imgf):
    """Integer counts (hist_a, hist_b, hist_f, joint_af, joint_bf) as int64 CPU tensors; a batch is ONE population, as in
    calc_prob / calc_joint_prob (metric.py:103-145)."""
    c = _hist(img1, img2, imgf)[2][0].to(torch.int64).sum(dim=0).cpu()
    return (c[0:256], c[256:512], c[512:768], c[768:768 + 65536].view(256, 256),
            c[768 + 65536:].view(256, 256))
