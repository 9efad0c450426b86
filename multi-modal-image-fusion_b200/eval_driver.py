"""Sharded evaluation driver: the pair loop of the reference's ``eval.py`` (eval.py:150-361) on the
B200 metric suite.

What the reference does per method: list ``img1_dir`` in natural order, read ``img1_dir/<name>``,
``img2_dir/<name>`` and ``imgf_dir/<i+1:02>.bmp`` as 8-bit grayscale (eval.py:176-187), run the 16
metrics of ``eval_metrics`` on float32 copies (eval.py:29-75, 189-206), then write one column per
metric — ``[header, mean, std, v_0, v_1, ...]`` — into an xlsx sheet (eval.py:231-305).

Here: pairs are sharded ``i % world == rank`` (one process per GPU), decoded by a thread pool (cv2
releases the GIL) while the GPU works, kept as uint8 until they are on the device (1 byte per pixel over
PCIe instead of 4; widening on the device is exact), batched by image shape into one launch per kernel
family, and the rows are gathered to rank 0, which applies the reference's aggregation verbatim
(including its std-over-the-list-that-already-holds-the-mean quirk) and writes the same sheet layout.
``natsort`` / ``openpyxl`` are not needed: a natural sort key and a minimal xlsx writer live below.
"""
import argparse
import concurrent.futures as cf
import os
import re
import time
import zipfile
from xml.sax.saxutils import escape

import numpy as np
import torch

from . import dist_utils as DU

HEADERS = ('SD', 'AG', 'SF', 'MSE', 'PSNR', 'CC', 'SCD', 'EN', 'CE', 'MI', 'Qabf', 'Nabf', 'Labf', 'SSIM', 'MSSSIM',
           'VIFF')   # eval.py:270-285, same order as dist_utils.METRIC_NAMES


def natural_key(name):
    """Sort key equivalent to natsort.natsorted for file names: digit runs compare as integers."""
    return [int(tok) if tok.isdigit() else tok.lower() for tok in re.split(r'(\d+)', name)]


def list_pairs(img1_dir, img2_dir, imgf_dir, fused_pattern='{index:0>2}.bmp'):
    """eval.py:176-180: the i-th file of img1_dir (natural order) pairs with the same name in img2_dir
    and with ``<i+1:02>.bmp`` in imgf_dir.  ``fused_pattern`` may also use ``{name}`` / ``{stem}``."""
    out = []
    for i, name in enumerate(sorted(os.listdir(img1_dir), key=natural_key)):
        stem = os.path.splitext(name)[0]
        fused = fused_pattern.format(index=i + 1, name=name, stem=stem)
        out.append((name, os.path.join(img1_dir, name), os.path.join(img2_dir, name), os.path.join(imgf_dir, fused)))
    return out


def decode_gray_u8(path):
    """cv2.imread(path, IMREAD_GRAYSCALE) (eval.py:182-187) -> (H, W) uint8; .npy arrays are accepted too."""
    if path.endswith('.npy'):
        img = np.load(path)
    else:
        import cv2
        img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
    if img is None:
        raise FileNotFoundError(f'cannot read image {path}')
    if img.dtype != np.uint8 or img.ndim != 2:
        raise ValueError(f'{path}: expected an 8-bit single-channel image, got {img.dtype} {img.shape}')
    return np.ascontiguousarray(img)


def _default_rows_u8(device):
    from .core.metric import eval_metrics_batch_u8

    def rows(a, b, f):          # pinned (n, H, W) uint8 host tensors -> (n, 16) float64 device tensor
        return eval_metrics_batch_u8(a.to(device, non_blocking=True), b.to(device, non_blocking=True),
                                     f.to(device, non_blocking=True))
    return rows


def evaluate_pairs(pairs, rank=0, world_size=1, device=None, compute_rows_u8=None, batch=8, workers=8, decode=decode_gray_u8,
                   group=None):
    """pairs: list of (name, path1, path2, pathf).  Returns (names, table) on rank 0 — table is the
    (n_pairs, 16) float64 array in pair order — and (names, None) on the other ranks."""
    if compute_rows_u8 is None:
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device())
        compute_rows_u8 = _default_rows_u8(device)
    pin = torch.cuda.is_available()
    mine = DU.shard_indices(len(pairs), rank, world_size)
    pending, idx, rows = {}, [], []

    def flush(shape):
        items = pending.pop(shape)
        stk = [torch.from_numpy(np.stack([it[k] for it in items])) for k in (1, 2, 3)]
        if pin:
            stk = [t.pin_memory() for t in stk]
        rows.append(compute_rows_u8(*stk))
        idx.extend(it[0] for it in items)

    def load(i):
        _, p1, p2, pf = pairs[i]
        a, b, f = decode(p1), decode(p2), decode(pf)
        if a.shape != b.shape or a.shape != f.shape:
            raise ValueError(f'pair {pairs[i][0]}: shapes differ {a.shape} {b.shape} {f.shape}')
        return i, a, b, f

    with cf.ThreadPoolExecutor(max_workers=max(1, workers)) as pool:
        for i, a, b, f in pool.map(load, mine):       # in order; decoding runs ahead of the GPU work
            pending.setdefault(a.shape, []).append((i, a, b, f))
            if len(pending[a.shape]) >= batch:
                flush(a.shape)
    for shape in list(pending):
        flush(shape)
    local = torch.cat([torch.as_tensor(r, dtype=torch.float64).cpu() for r in rows]) if rows else \
        torch.empty(0, len(DU.METRIC_NAMES), dtype=torch.float64)
    table = DU.gather_rows(local, idx, len(pairs), rank, world_size, device=device, group=group)
    names = [p[0] for p in pairs]
    return names, (table.numpy() if table is not None else None)


def method_sheet(names, table):
    """The 'method' sheet of eval.py:268-305 as a list of 17 columns: column 0 = ['', 'mean', 'std', names...],
    column k = [HEADERS[k-1], mean, std, values...] with eval.py:231-266's mean / std."""
    cols = DU.aggregate_columns(table)
    sheet = [[''] + ['mean', 'std'] + list(names)]
    for hdr, key in zip(HEADERS, DU.METRIC_NAMES):
        sheet.append([hdr] + cols[key])
    return sheet


def write_csv(path, sheet):
    nrow = max(len(c) for c in sheet)
    with open(path, 'w') as fh:
        for r in range(nrow):
            fh.write(','.join('' if r >= len(c) else (repr(float(c[r])) if isinstance(c[r], (float, np.floating)) else str(c[r]))
                              for c in sheet) + '\n')


def _col_letter(k):
    s = ''
    k += 1
    while k:
        k, r = divmod(k - 1, 26)
        s = chr(65 + r) + s
    return s


def write_xlsx(path, sheets):
    """Minimal xlsx (zip of XML parts; inline strings, no styles): sheets = {title: list of columns}.
    Cell (row r, column k) holds sheets[title][k][r] like write_excel(file, sheet, column, data) of eval.py:76-96."""
    def sheet_xml(cols):
        nrow = max((len(c) for c in cols), default=0)
        out = ['<?xml version="1.0" encoding="UTF-8" standalone="yes"?>',
               '<worksheet xmlns="http://schemas.openxmlformats.org/spreadsheetml/2006/main"><sheetData>']
        for r in range(nrow):
            out.append(f'<row r="{r + 1}">')
            for k, c in enumerate(cols):
                if r >= len(c) or c[r] is None:
                    continue
                ref = f'{_col_letter(k)}{r + 1}'
                v = c[r]
                if isinstance(v, (int, float, np.integer, np.floating)) and np.isfinite(v):
                    out.append(f'<c r="{ref}"><v>{float(v)!r}</v></c>')
                else:
                    out.append(f'<c r="{ref}" t="inlineStr"><is><t>{escape(str(v))}</t></is></c>')
            out.append('</row>')
        out.append('</sheetData></worksheet>')
        return ''.join(out)

    titles = list(sheets)
    with zipfile.ZipFile(path, 'w', zipfile.ZIP_DEFLATED) as z:
        z.writestr('[Content_Types].xml',
                   '<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
                   '<Types xmlns="http://schemas.openxmlformats.org/package/2006/content-types">'
                   '<Default Extension="rels" ContentType="application/vnd.openxmlformats-package.relationships+xml"/>'
                   '<Default Extension="xml" ContentType="application/xml"/>'
                   '<Override PartName="/xl/workbook.xml" ContentType="application/vnd.openxmlformats-officedocument.spreadsheetml.sheet.main+xml"/>'
                   + ''.join(f'<Override PartName="/xl/worksheets/sheet{i + 1}.xml" ContentType="application/vnd.openxmlformats-officedocument.spreadsheetml.worksheet+xml"/>'
                             for i in range(len(titles))) + '</Types>')
        z.writestr('_rels/.rels',
                   '<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
                   '<Relationships xmlns="http://schemas.openxmlformats.org/package/2006/relationships">'
                   '<Relationship Id="rId1" Type="http://schemas.openxmlformats.org/officeDocument/2006/relationships/officeDocument" Target="xl/workbook.xml"/>'
                   '</Relationships>')
        z.writestr('xl/workbook.xml',
                   '<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
                   '<workbook xmlns="http://schemas.openxmlformats.org/spreadsheetml/2006/main" '
                   'xmlns:r="http://schemas.openxmlformats.org/officeDocument/2006/relationships"><sheets>'
                   + ''.join(f'<sheet name="{escape(t)}" sheetId="{i + 1}" r:id="rId{i + 1}"/>' for i, t in enumerate(titles))
                   + '</sheets></workbook>')
        z.writestr('xl/_rels/workbook.xml.rels',
                   '<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
                   '<Relationships xmlns="http://schemas.openxmlformats.org/package/2006/relationships">'
                   + ''.join(f'<Relationship Id="rId{i + 1}" Type="http://schemas.openxmlformats.org/officeDocument/2006/relationships/worksheet" Target="worksheets/sheet{i + 1}.xml"/>'
                             for i in range(len(titles))) + '</Relationships>')
        for i, t in enumerate(titles):
            z.writestr(f'xl/worksheets/sheet{i + 1}.xml', sheet_xml(sheets[t]))


def main(argv=None):
    ap = argparse.ArgumentParser(description='Sharded eval.py pair loop on the B200 metric suite')
    ap.add_argument('--img1-dir', required=True)
    ap.add_argument('--img2-dir', required=True)
    ap.add_argument('--imgf-dir', required=True)
    ap.add_argument('--fused-pattern', default='{index:0>2}.bmp', help="eval.py:180 default; may use {name} / {stem}")
    ap.add_argument('--method', default='DeepFuse', help='sheet title (eval.py: method_names[0])')
    ap.add_argument('--out', default='metrics.xlsx', help='.xlsx or .csv, written by rank 0')
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--workers', type=int, default=8)
    args = ap.parse_args(argv)
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    pairs = list_pairs(args.img1_dir, args.img2_dir, args.imgf_dir, args.fused_pattern)
    t0 = time.time()
    names, table = evaluate_pairs(pairs, rank, world, dev, batch=args.batch, workers=args.workers)
    if rank == 0:
        sheet = method_sheet(names, table)
        if args.out.endswith('.csv'):
            write_csv(args.out, sheet)
        else:
            write_xlsx(args.out, {args.method: sheet})
        print(f'evaluating {args.method} done: {len(pairs)} pairs on {world} GPU(s), cost {time.time() - t0:.3f}s -> {args.out}')
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
