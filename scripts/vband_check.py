"""grad_value kernels side by side: dense-band (default for group_channels 16) against the scatter-tile tcgen05 kernel
(DCNV3_VALUE=vmma) and the fp64 direct oracle.  Parity on small / ragged / cfg2 shapes, then timings.
    python scripts/vband_check.py [--time-only]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tests/golden')
from oracle import dcnv3_oracle as orc
import DCNv3


def run(which, dv, do_, dm, dg, geom):
    os.environ['DCNV3_VALUE'] = 'band' if which == 'vband' else 'vmma'
    grads = DCNv3.dcnv3_backward(dv, do_, dm, *geom, dg, 256)
    torch.cuda.synchronize()
    return grads


def errs(a, w):
    a = a.double().cpu().numpy(); rms = float(np.sqrt(np.mean(w ** 2)))
    d = np.abs(a - w)
    return dict(max_abs=float(d.max()), max_over_rms=float(d.max() / rms), bad=float(np.mean(d > 1e-2 * np.abs(w) + 1e-2 * rms)))


shapes = [(2, 20, 24, 8, 1.0, 1.0), (1, 19, 21, 8, 1.0, 2.0), (1, 8, 8, 8, 1.0, 0.0), (2, 37, 50, 16, 0.75, 1.0), (1, 80, 80, 16, 1.0, 1.0),
          (16, 80, 80, 16, 1.0, 1.0)]
if '--time-only' in sys.argv:
    shapes = shapes[-1:]
for dt in (torch.bfloat16, torch.float16):
    for (N, H, W, G, sigma, ostd) in shapes:
        gc = 16
        geom = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, sigma)
        g = torch.Generator().manual_seed(N * 1000 + H)
        v = torch.randn(N, H, W, G * gc, generator=g); o = torch.randn(N, H, W, G * 18, generator=g) * ostd
        m = torch.softmax(torch.randn(N, H, W, G, 9, generator=g), -1).reshape(N, H, W, -1); go = torch.randn(N, H, W, G * gc, generator=g)
        arrs = [t.to(dt) for t in (v, o, m, go)]
        dv, do_, dm, dg = (t.cuda() for t in arrs)
        res = {}
        for which in ('vband', 'vmma'):
            res[which] = run(which, dv, do_, dm, dg, geom)
        line = f"{str(dt)[6:]} N{N} {H}x{W} G{G} s{sigma} o{ostd}:"
        if N * H * W <= 2 * 80 * 80 and '--time-only' not in sys.argv:
            a64 = [t.double().numpy() for t in arrs]
            want = orc.direct_backward(*a64, *geom)
            for which in ('vband', 'vmma'):
                line += f"  {which} gv {errs(res[which][0], want[0])}"
            line += f"  go {errs(res['vband'][1], want[1])['bad']:.2e} gm {errs(res['vband'][2], want[2])['bad']:.2e}"
        d = (res['vband'][0].float() - res['vmma'][0].float()).abs().max().item()
        line += f"  |vband - vmma|max {d:.4g}"
        # timing of the backward pass, both forms
        for which in ('vband', 'vmma'):
            for _ in range(3): run(which, dv, do_, dm, dg, geom)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): DCNv3.dcnv3_backward(dv, do_, dm, *geom, dg, 256)
            e1.record(); torch.cuda.synchronize()
            line += f"  bwd[{which}] {e0.elapsed_time(e1) * 50:.1f} us"
        print(line, flush=True)
os.environ.pop('DCNV3_VALUE', None)
