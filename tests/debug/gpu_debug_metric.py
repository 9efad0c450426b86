import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'tests', 'golden')]
import numpy as np, torch
import cases, gates
import mmif_b200
from mmif_b200.core import metric as MM
from oracle import fusion_metric as OM
MG = np.load(cases.HERE + '/metric_golden.npz')
for name in (sys.argv[1:] or cases.METRIC_CASES):
    a, b, f = (torch.from_numpy(x).cuda() for x in cases.metric_case(name))
    row = MM.eval_metrics(a, b, f)
    r32, r64 = MG[f'{name}/f32/metrics'], MG[f'{name}/f64/metrics']
    print(name)
    for k, nm in enumerate(OM.METRIC_NAMES):
        ok = gates.scalar_ok(row[nm], r32[k], r64[k])
        print(f'   {nm:7s} new {row[nm]: .9e} ref32 {r32[k]: .9e} ref64 {r64[k]: .9e} rel32 {abs(row[nm]-r32[k])/max(abs(r32[k]),1e-300):.2e} rel64 {abs(row[nm]-r64[k])/max(abs(r64[k]),1e-300):.2e} ref32-64 {abs(r32[k]-r64[k])/max(abs(r64[k]),1e-300):.2e} {"" if ok else "FAIL"}')
    ms = MM._msssim2(a, b, f, 11, 255.0, False)[0].cpu().numpy()
    vf = MM._viff3(a, b, f)[0].cpu().numpy()
    print('   msssim levels', ms[2:].reshape(5, 4))
    print('   viff scales', vf[2:].reshape(4, 6))
