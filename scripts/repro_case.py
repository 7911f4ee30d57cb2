"""Run one parity case of tests/golden/cases.py through the CUDA path: python scripts/repro_case.py NAME [spread]"""
import sys
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tests/golden')
import cases
from test_dcnv3_gpu import _TILE_CASES, run_cuda, rounded
from oracle import dcnv3_oracle as orc
name = sys.argv[1]; spread = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
case = next(c for c in _TILE_CASES if c.name == name)
v, o, m, g = cases.make_inputs(case)
arrs = rounded((v, o * spread, m, g), torch.bfloat16)
got = run_cuda(arrs, case.geom, dtype=torch.bfloat16)
out = orc.direct_forward(*arrs[:3], *case.geom)
want = (out, *orc.direct_backward(*arrs, *case.geom))
for nm, a, w in zip(("out", "gv", "go", "gm"), got, want):
    rms = float(np.sqrt(np.mean(w ** 2)))
    bad = np.abs(a - w) > 1e-2 * np.abs(w) + 1e-2 * rms
    print(nm, 'bad frac', bad.mean(), 'max', np.abs(a - w).max(), 'rms', rms)
    if nm == 'gv' and bad.any():
        idx = np.argwhere(bad)
        print('  bad n', np.unique(idx[:, 0]), 'rows', idx[:, 1].min(), idx[:, 1].max(), 'cols', idx[:, 2].min(), idx[:, 2].max(), 'ch', np.unique(idx[:, 3] // 16))
        rows = np.bincount(idx[:, 1], minlength=a.shape[1]); print('  per-row bad', rows.tolist())
        cols = np.bincount(idx[:, 2], minlength=a.shape[2]); print('  per-col bad', cols.tolist())
