"""CPU: pin the oracle (oracle/) against golden vectors produced by the REAL reference
(tests/golden/make_golden.py ran dcnv3_core_pytorch, models/ops_dcnv3/functions/dcnv3_func.py:147-188).

Tolerances
* core_gridsample is the same algorithm on the same torch build -> bit-tight (<= 1e-6 abs fp32,
  1e-12 fp64; autograd summation order is identical too).
* the direct pixel-space form differs from the reference's normalise->un-normalise fp32 coordinate
  path: vs the fp64 reference run the residual is the reference's own fp32 coordinate
  rounding, bounded here at 2e-5 x max|reference| (floor: the data scale, 1 or 0.01).
"""
import numpy as np
import pytest
import torch

import cases
from helpers import check_inputs_unchanged, golden, max_abs, view_like_golden
from oracle import dcnv3_oracle as orc

WHAT = ("out", "gv", "go", "gm")


@pytest.mark.parametrize("case", cases.ALL, ids=lambda c: c.name)
@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_gridsample_restatement_matches_reference(case, dt):
    arrs = check_inputs_unchanged(case)
    tdt = torch.float64 if dt == "f64" else torch.float32
    v, o, m, g = (torch.from_numpy(a).to(tdt) for a in arrs)
    got = orc.gridsample_fwd_bwd(v, o, m, g, *case.geom)
    tol = 1e-12 if dt == "f64" else 1e-6
    for name, arr in zip(WHAT, got):
        kind, want, s = golden(case.name, dt, name)
        a = arr.numpy()
        assert max_abs(view_like_golden(kind, a), want) <= tol * max(1.0, float(np.abs(want).max())), name
        assert abs(float(a.astype(np.float64).sum()) - s) <= 1e-4 * max(1.0, abs(s)), name


@pytest.mark.parametrize("case", cases.ALL, ids=lambda c: c.name)
def test_direct_f64_matches_reference_f64(case):
    arrs = check_inputs_unchanged(case)
    v, o, m, g = arrs
    out = orc.direct_forward(v, o, m, *case.geom)
    gv, go, gm = orc.direct_backward(v, o, m, g, *case.geom)
    scale = 0.01 if case.dist == "reftest" else 1.0
    for name, a in zip(WHAT, (out, gv, go, gm)):
        kind, want, _ = golden(case.name, "f64", name)
        # grad wrt offset is discontinuous where a sample crosses a pixel boundary; the reference's
        # fp32 coordinates may floor() differently there -> compare robustly (99.9th percentile)
        diff = np.abs(view_like_golden(kind, a).astype(np.float64) - want)
        bound = 2e-5 * max(scale, float(np.abs(want).max()))
        if name == "go":
            assert np.quantile(diff, 0.999) <= bound, (name, float(diff.max()))
        else:
            assert diff.max() <= bound, (name, float(diff.max()))


@pytest.mark.parametrize("case", [cases.REFTEST_FWD, cases.SWEEP[0], cases.SWEEP[3]],
                         ids=lambda c: c.name)
def test_direct_f32_close_to_direct_f64(case):
    v, o, m, g = cases.make_inputs(case)
    f32 = lambda a: a.astype(np.float32)
    out64 = orc.direct_forward(v, o, m, *case.geom)
    out32 = orc.direct_forward(f32(v), f32(o), f32(m), *case.geom, dtype=np.float32)
    assert out32.dtype == np.float32
    assert max_abs(out32, out64) <= 5e-5 * max(1.0, float(np.abs(out64).max()))


def test_kat_average_pool():
    """KAT-1: zero offsets + uniform mask == 3x3 average pool, count_include_pad."""
    rng = np.random.default_rng(0)
    N, H, W, G, gc = 2, 7, 9, 2, 4
    v = rng.standard_normal((N, H, W, G * gc))
    o = np.zeros((N, H, W, G * 9 * 2))
    m = np.full((N, H, W, G * 9), 1.0 / 9)
    out = orc.direct_forward(v, o, m, 3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    want = torch.nn.functional.avg_pool2d(torch.from_numpy(v).permute(0, 3, 1, 2), 3, 1, 1,
                                          count_include_pad=True).permute(0, 2, 3, 1).numpy()
    assert max_abs(out, want) <= 1e-12
    out_gs = orc.core_gridsample(torch.from_numpy(v), torch.from_numpy(o), torch.from_numpy(m),
                                 3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0).numpy()
    assert max_abs(out_gs, want) <= 5e-6


def test_kat_one_hot_integer_shift():
    """KAT-2: one-hot mask on point p with integer offset (dx,dy) copies
    value[n, y+(p%K)-1+dy, x+(p//K)-1+dx] -- pins p = i_x*K + j_y and the (dx,dy) interleave."""
    rng = np.random.default_rng(1)
    N, H, W, G, gc, K = 1, 6, 8, 1, 3, 3
    v = rng.standard_normal((N, H, W, G * gc))
    for p, dx, dy in [(0, 0, 0), (5, 1, -1), (7, -2, 1), (2, 0, 2)]:
        o = np.zeros((N, H, W, G * 9, 2)); o[..., p, 0] = dx; o[..., p, 1] = dy
        m = np.zeros((N, H, W, G * 9)); m[..., p] = 1.0
        out = orc.direct_forward(v, o.reshape(N, H, W, -1), m, K, K, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
        want = np.zeros_like(v)
        for y in range(H):
            for x in range(W):
                yy, xx = y + (p % K) - 1 + dy, x + (p // K) - 1 + dx
                if 0 <= yy < H and 0 <= xx < W:
                    want[0, y, x] = v[0, yy, xx]
        assert max_abs(out, want) <= 1e-12, (p, dx, dy)


def test_direct_backward_matches_autograd_of_direct_formula():
    """Analytic grads (cuh:82-147) vs central finite differences of the direct forward."""
    c = cases.Case("fd", N=1, H=5, W=6, G=1, gc=2, sigma=1.3, seed=55)
    v, o, m, g = cases.make_inputs(c)
    gv, go, gm = orc.direct_backward(v, o, m, g, *c.geom)
    f = lambda vv, oo, mm: float((orc.direct_forward(vv, oo, mm, *c.geom) * g).sum())
    rng = np.random.default_rng(3)
    eps = 1e-6
    for arr, grad, idx_n in ((v, gv, 12), (m, gm, 12), (o, go, 24)):
        flat = arr.reshape(-1)
        for i in rng.choice(flat.size, idx_n, replace=False):
            old = flat[i]
            flat[i] = old + eps; up = f(v, o, m)
            flat[i] = old - eps; dn = f(v, o, m)
            flat[i] = old
            fd = (up - dn) / (2 * eps)
            assert abs(fd - grad.reshape(-1)[i]) <= 1e-5 * max(1.0, abs(fd)), (i, fd, grad.reshape(-1)[i])
