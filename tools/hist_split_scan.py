"""Scan of hist_kernel's pixel-split factor S (MMIF_HIST_SPLIT) on the two BASELINE metric shapes: times mmif_hist alone and
the whole suite, one subprocess per setting (the library reads the variable once).  Tuning aid for launch_hist's cost model."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys
sys.path.insert(0, %r)
import torch, mmif_b200
from mmif_b200.core import metric as MM
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
for (n, h, w) in ((21, 480, 640), (32, 1024, 1224), (4, 1024, 1224), (1, 1024, 1224)):
    g = torch.Generator(device='cuda').manual_seed(7)
    a = torch.randint(0, 256, (n, 1, h, w), device='cuda', generator=g).float()
    b = torch.randint(0, 256, (n, 1, h, w), device='cuda', generator=g).float()
    f = torch.floor((a + b) / 2)
    def hist():
        MM._memo.items.clear()
        MM.hist_raw(a, b, f)
    print('S=%%s %%dx%%dx%%d: hist %%.1f us  suite %%.1f us' %% (os.environ.get('MMIF_HIST_SPLIT', 'model'), n, h, w, t(hist), t(lambda: MM.eval_metrics_batch(a, b, f))))
''' % ROOT
for s in (None, '1', '2', '3', '4', '6'):
    env = dict(os.environ)
    if s is None:
        env.pop('MMIF_HIST_SPLIT', None)
    else:
        env['MMIF_HIST_SPLIT'] = s
    subprocess.run([sys.executable, '-c', CHILD], env=env, check=False)
