"""Sharded evaluation driver (eval.py:150-361 pair loop): natural order, pairing rule, sheet layout, xlsx writer,
world_size-2 gloo run == single-process run (CPU, injected row function), and the real GPU suite on decoded files."""
import os
import re
import socket
import zipfile

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _drv():
    import mmif_b200  # noqa: F401
    from mmif_b200 import eval_driver as ED
    return ED


def _make_dataset(tmp, n=7, seed=0):
    import cv2
    rng = np.random.default_rng(seed)
    d1, d2, df = (os.path.join(tmp, k) for k in ('vis', 'ir', 'fused'))
    for d in (d1, d2, df):
        os.makedirs(d, exist_ok=True)
    names = [f'{i}.png' for i in range(1, n + 1)]             # 1.png ... 10.png: natural order != lexicographic
    arrays = {}
    for k, name in enumerate(sorted(names, key=lambda s: int(s.split('.')[0]))):
        shape = (96, 128) if k % 3 else (80, 144)
        a = rng.integers(0, 256, shape, dtype=np.uint8)
        b = rng.integers(0, 256, shape, dtype=np.uint8)
        f = np.maximum(a, b)
        cv2.imwrite(os.path.join(d1, name), a)
        cv2.imwrite(os.path.join(d2, name), b)
        cv2.imwrite(os.path.join(df, f'{k + 1:0>2}.bmp'), f)
        arrays[name] = (a, b, f)
    return d1, d2, df, arrays


def _fake_rows(a, b, f):       # (n,H,W) uint8 -> (n,16): depends on all three images, cheap, exact
    base = a.double().mean(dim=(1, 2)) + 2 * b.double().mean(dim=(1, 2)) + 3 * f.double().mean(dim=(1, 2))
    return torch.stack([base + k for k in range(16)], dim=1)


def test_natural_order_and_pairing(tmp_path):
    ED = _drv()
    assert sorted(['10.png', '9.png', '1.png', 'a2.png', 'a10.png'], key=ED.natural_key) == ['1.png', '9.png', '10.png', 'a2.png', 'a10.png']
    d1, d2, df, _ = _make_dataset(str(tmp_path), n=11)
    pairs = ED.list_pairs(d1, d2, df)
    assert [p[0] for p in pairs][:3] == ['1.png', '2.png', '3.png'] and pairs[9][0] == '10.png'
    assert pairs[9][3].endswith(os.path.join('fused', '10.bmp')) and pairs[0][3].endswith('01.bmp')     # eval.py:180


def test_sheet_layout_and_xlsx_roundtrip(tmp_path):
    ED = _drv()
    d1, d2, df, arrays = _make_dataset(str(tmp_path), n=7)
    names, table = ED.evaluate_pairs(ED.list_pairs(d1, d2, df), compute_rows_u8=_fake_rows, batch=2, workers=2)
    assert names == [f'{i}.png' for i in range(1, 8)]
    for i, nm in enumerate(names):
        a, b, f = arrays[nm]
        assert table[i, 0] == a.astype(np.float64).mean() + 2 * b.astype(np.float64).mean() + 3 * f.astype(np.float64).mean()
    sheet = ED.method_sheet(names, table)
    assert len(sheet) == 17 and sheet[0][:3] == ['', 'mean', 'std'] and sheet[1][0] == 'SD' and sheet[16][0] == 'VIFF'
    vals = list(table[:, 0])
    mean = np.mean(vals)
    assert sheet[1][1] == mean and sheet[1][2] == np.std([mean] + vals)      # eval.py:231-266: std over the list holding the mean
    assert sheet[1][3:] == vals
    out = str(tmp_path / 'm.xlsx')
    ED.write_xlsx(out, {'DeepFuse': sheet})
    with zipfile.ZipFile(out) as z:
        assert {'[Content_Types].xml', 'xl/workbook.xml', 'xl/worksheets/sheet1.xml'} <= set(z.namelist())
        xml = z.read('xl/worksheets/sheet1.xml').decode()
        assert 'name="DeepFuse"' in z.read('xl/workbook.xml').decode()
    assert re.search(r'<c r="B1" t="inlineStr"><is><t>SD</t>', xml) and re.search(r'<c r="A4" t="inlineStr"><is><t>1.png</t>', xml)
    b2 = float(re.search(r'<c r="B2"><v>([^<]+)</v>', xml).group(1))
    assert b2 == mean
    ED.write_csv(str(tmp_path / 'm.csv'), sheet)
    assert open(str(tmp_path / 'm.csv')).readline().startswith(',SD,AG,SF')


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp, q):
    import sys
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import mmif_b200  # noqa: F401
    from mmif_b200 import eval_driver as ED
    pairs = ED.list_pairs(os.path.join(tmp, 'vis'), os.path.join(tmp, 'ir'), os.path.join(tmp, 'fused'))
    names, table = ED.evaluate_pairs(pairs, rank, world, compute_rows_u8=_fake_rows, batch=2, workers=2)
    q.put((rank, None if table is None else table.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_equals_single_process(tmp_path):
    ED = _drv()
    tmp = str(tmp_path)
    d1, d2, df, _ = _make_dataset(tmp, n=7)
    _, single = ED.evaluate_pairs(ED.list_pairs(d1, d2, df), compute_rows_u8=_fake_rows, batch=3, workers=1)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, tmp, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[1] is None and got[0] == single.tolist()


@pytest.mark.gpu
def test_driver_on_gpu_equals_batched_suite(tmp_path):
    ED = _drv()
    from mmif_b200.core import metric as MM
    d1, d2, df, arrays = _make_dataset(str(tmp_path), n=6, seed=3)
    names, table = ED.evaluate_pairs(ED.list_pairs(d1, d2, df), device=torch.device('cuda', 0), batch=4, workers=4)
    for i, nm in enumerate(names):
        a, b, f = (torch.from_numpy(x).float()[None, None].cuda() for x in arrays[nm])
        ref = MM.eval_metrics_batch(a, b, f)[0].cpu().numpy()
        np.testing.assert_allclose(table[i], ref, rtol=1e-7, atol=1e-12)       # row blocking differs with the batch size
    out = str(tmp_path / 'metrics.xlsx')
    ED.write_xlsx(out, {'MyFusion': ED.method_sheet(names, table)})
    assert zipfile.is_zipfile(out)
