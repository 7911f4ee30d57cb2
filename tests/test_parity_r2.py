"""Round-2 parity additions (VERDICT round 1, items 3 and the advisor's findings):
  * backward known-answer cases on the LATTICE (zero / integer offsets, pad 1: the outer taps of every border pixel sit
    at loc == -1 exactly, where the reference's range test `loc > -1` (dcnv3_im2col_cuda.cuh:334) zeroes the point) for
    every backward kernel family, with an absolute error cap;
  * BASELINE configs[4] (192 x 192, C = 256, G in {8, 16, 32}) against the C oracle, bf16 and fp16;
  * the deterministic backward against the ORACLE (round 1 compared it with the default path only);
  * containment of non-finite inputs (what the tensor-core value kernel may and may not poison);
  * the same for the reference's own layer defaults (C = 64, G = 4; offset_scale = 2).
All through DCNv3Function -> DCNv3 shim -> C ABI."""
import numpy as np
import pytest
import torch

import cases
from helpers import allclose_frac, max_abs
from test_dcnv3_gpu import TDT, WHAT, rounded, run_cuda

pytestmark = pytest.mark.gpu


def _oracle(arrs, geom):
    from oracle import dcnv3_oracle as orc
    return (orc.direct_forward(*arrs[:3], *geom), *orc.direct_backward(*arrs, *geom))


def _check_16bit(got, want, dt, go_frac=5e-4, frac=3e-4, cap=8e-2):
    """1e-2 relative (north_star) with a floor of 1e-2 x RMS; at most `frac` of the elements outside (grad_offset:
    `go_frac`, floor() flips at pixel boundaries); and NO element further off than cap x max(RMS, |w|) -- except
    grad_offset, whose flipped elements are legitimately a different branch of a discontinuous function."""
    for name, a, w in zip(WHAT, got, want):
        rms = float(np.sqrt(np.mean(w ** 2))) + 1e-30
        f = allclose_frac(a, w, rtol=1e-2, atol=1e-2 * rms)
        assert f <= (go_frac if name == "go" else frac), (name, f, max_abs(a, w), rms)
        if name != "go":
            worst = float(np.max(np.abs(a - w) / np.maximum(np.abs(w), rms)))
            assert worst <= (cap if dt != "f16" else cap / 4), (name, worst)


# ----------------------------------------------------------------------------- lattice known-answer cases
def _lattice_inputs(c, mode, seed):
    rng = np.random.default_rng(seed)
    v, _, m, g = cases.make_inputs(c)
    shape = (c.N, c.Ho, c.Wo, c.G * c.P * 2)
    if mode == "zero":
        o = np.zeros(shape)
    elif mode == "int":
        o = rng.integers(-2, 3, size=shape).astype(np.float64)
    else:  # half-integers and integers mixed: taps on pixel centres and exactly between pixels
        o = rng.integers(-4, 5, size=shape).astype(np.float64) * 0.5
    return v, o, m, g


@pytest.mark.parametrize("bwd", ["default", "strip", "mma", "scatter"])
@pytest.mark.parametrize("mode", ["zero", "int", "half"])
@pytest.mark.parametrize("dt", ["bf16", "f32"])
def test_backward_lattice_offsets_all_kernel_families(dt, mode, bwd, monkeypatch):
    """The state of every layer at initialisation (zero offsets) and exact-lattice offsets: floor() sits ON its
    discontinuity, and border pixels have taps at loc == -1 where grad_offset / grad_mask / grad_value must be exactly
    zero.  The fp64 oracle decides; no allowance for grad_offset flips here -- on the lattice the coordinates are exact
    in every association, so all kernels must land on the same side."""
    if bwd != "default":
        if dt == "f32" and bwd != "scatter":
            pytest.skip("fp32 I/O has one backward kernel family")
        monkeypatch.setenv("DCNV3_BWD", bwd)
    c = cases.Case("lattice", N=2, H=21, W=26, G=8, gc=16, seed=91)
    arrs = rounded(_lattice_inputs(c, mode, 7), TDT[dt])
    got = run_cuda(arrs, c.geom, dtype=TDT[dt])
    want = _oracle(arrs, c.geom)
    if dt == "f32":
        for name, a, w in zip(WHAT, got, want):
            assert allclose_frac(a, w, rtol=1e-5, atol=2e-6 * (1.0 + float(np.abs(w).max()))) == 0.0, (name, max_abs(a, w))
    else:
        _check_16bit(got, want, dt, go_frac=0.0, frac=3e-4)
    # the points the reference's range test rejects have exactly zero gradients
    go, gm = got[2], got[3]
    o = arrs[1].reshape(c.N, c.Ho, c.Wo, c.G, c.P, 2)
    ys, xs = np.arange(c.Ho)[None, :, None, None, None], np.arange(c.Wo)[None, None, :, None, None]
    p = np.arange(c.P)[None, None, None, None, :]
    loc_w = xs - 1 + (p // 3) + o[..., 0]
    loc_h = ys - 1 + (p % 3) + o[..., 1]
    rejected = ~((loc_h > -1) & (loc_w > -1) & (loc_h < c.H) & (loc_w < c.W))
    assert rejected.any()
    assert np.all(go.reshape(o.shape)[rejected] == 0.0)
    assert np.all(gm.reshape(rejected.shape)[rejected] == 0.0)


# ----------------------------------------------------------------------------- BASELINE configs[4]
@pytest.mark.parametrize("dt", ["bf16", "f16"])
@pytest.mark.parametrize("G", [8, 16, 32])
def test_cfg5_192x192_against_direct_oracle(G, dt):
    """BASELINE.json configs[4]: 1536 x 1536 input -> DCNv3 at 192 x 192, C = 256, group 8 / 16 / 32 (group_channels
    32 / 16 / 8), one image against the C oracle in fp64 (2.7 / 5.3 / 10.6 M sampled points)."""
    c = cases.Case(f"cfg5_g{G}", N=1, H=192, W=192, G=G, gc=256 // G, seed=500 + G)
    v, o, m, g = cases.make_inputs(c)
    arrs = rounded((v, o / 1.5, m, g), TDT[dt])            # N(0, 1)-pixel offsets
    got = run_cuda(arrs, c.geom, dtype=TDT[dt])
    _check_16bit(got, _oracle(arrs, c.geom), dt, go_frac=2e-4)


@pytest.mark.parametrize("G", [8, 16, 32])
def test_cfg5_batch4_properties(G):
    """The same shapes at N = 4 (the batch the bench quotes for configs[4]) through size-independent properties:
    the adjoint identities <f(v), g> = <v, grad_v> = <m, grad_m>."""
    from yolo_somi_b200.ops_dcnv3.functions import DCNv3Function
    gen = torch.Generator(device="cuda").manual_seed(G)
    N, H, W, gc = 4, 192, 192, 256 // G
    dt = torch.bfloat16
    v = torch.randn(N, H, W, 256, device="cuda", generator=gen).to(dt).requires_grad_(True)
    o = torch.randn(N, H, W, G * 18, device="cuda", generator=gen).to(dt)
    m = torch.softmax(torch.randn(N, H, W, G, 9, device="cuda", generator=gen), -1).reshape(N, H, W, -1).to(dt).requires_grad_(True)
    g = torch.randn(N, H, W, 256, device="cuda", generator=gen).to(dt)
    out = DCNv3Function.apply(v, o, m, 3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0, 256)
    out.backward(g)
    dot_out = float((out.double() * g.double()).sum())
    ref = float((out.double().abs() * g.double().abs()).sum())
    assert abs(dot_out - float((v.grad.double() * v.double()).sum())) <= 2e-3 * ref
    assert abs(dot_out - float((m.grad.double() * m.double()).sum())) <= 2e-3 * ref


# ----------------------------------------------------------------------------- deterministic mode vs the oracle
@pytest.mark.parametrize("dt", ["f32", "bf16", "f16"])
def test_deterministic_backward_against_the_oracle(dt, monkeypatch):
    """DCNV3_DETERMINISTIC=1 (64-bit fixed-point accumulation of grad_value, every coefficient in fp32) against the
    fp64 oracle -- at the fp32 tolerance for fp32 I/O and with fp32-accurate coefficients for 16-bit I/O (the errors
    left are the roundings of the outputs themselves: no violation allowance for grad_value)."""
    monkeypatch.setenv("DCNV3_DETERMINISTIC", "1")
    c = cases.Case("det_oracle", N=2, H=33, W=29, G=8, gc=16, seed=78)
    arrs = rounded(cases.make_inputs(c), TDT[dt])
    runs = [run_cuda(arrs, c.geom, dtype=TDT[dt]) for _ in range(2)]
    for a, b in zip(*runs):
        assert np.array_equal(a, b)
    want = _oracle(arrs, c.geom)
    if dt == "f32":
        for name, a, w in zip(WHAT, runs[0], want):
            lim = 1e-3 if name == "go" else 0.0
            assert allclose_frac(a, w, rtol=1e-5, atol=1e-5 * (1.0 + float(np.sqrt(np.mean(w ** 2))))) <= lim, (name, max_abs(a, w))
    else:
        eps = 2.0 ** -8 if dt == "bf16" else 2.0 ** -11
        gv, w = runs[0][1], want[1]
        assert float(np.max(np.abs(gv - w) / np.maximum(np.abs(w), np.sqrt(np.mean(w ** 2))))) <= 1.01 * eps + 1e-6
        _check_16bit(runs[0], want, dt)


# ----------------------------------------------------------------------------- non-finite containment
@pytest.mark.parametrize("dt", ["bf16", "f32"])
def test_nonfinite_value_poisons_only_what_samples_it(dt):
    """One NaN in `value`: `out`, grad_offset and grad_mask are non-finite exactly where the oracle's are (the gather
    kernels multiply a corner by its weight only where a point samples it -- clamped reads keep the reference's
    NaN propagation, dcnv3_common.cuh make_clamped_tap); grad_value does not depend on value and stays finite."""
    c = cases.Case("nan_value", N=1, H=24, W=24, G=8, gc=16, seed=92)
    v, o, m, g = cases.make_inputs(c)
    v[0, 11, 13, 35] = np.nan
    arrs = rounded((v, o, m, g), TDT[dt])
    got = run_cuda(arrs, c.geom, dtype=TDT[dt])
    want = _oracle(arrs, c.geom)
    assert np.isfinite(got[1]).all()
    for name, a, w in ((WHAT[0], got[0], want[0]), (WHAT[2], got[2], want[2]), (WHAT[3], got[3], want[3])):
        bad_a, bad_w = ~np.isfinite(a), ~np.isfinite(w)
        assert bad_w.any(), name
        # the oracle multiplies by zero weights explicitly (0 * NaN): it may mark MORE than the sampled points;
        # the kernel must never mark an element the oracle leaves finite, and must mark every element whose
        # non-zero-weight corner is the NaN (checked through the finite remainder being correct)
        assert not (bad_a & ~bad_w).any(), name
        ok = ~bad_w & ~bad_a
        rms = float(np.sqrt(np.mean(w[ok] ** 2)))
        tol = 1e-5 if dt == "f32" else 1e-2
        assert allclose_frac(a[ok], w[ok], rtol=tol, atol=tol * 10 * rms) <= (2e-3 if name == "go" else 1e-3), name


def test_nonfinite_grad_out_is_contained_by_the_value_kernel():
    """One Inf in grad_out (16-bit default backward, tcgen05 value kernel).  The reference adds w * m * Inf only at the
    four corners of the nine points of that (pixel, group).  The dense product A x grad_out also forms 0 x Inf = NaN
    for every cell of the product's band (a documented deviation) -- this test pins how far that
    reaches: the poisoned pixel's own group channel only, at most 16 cells around it in x and 16 in y, and nothing in
    grad_offset / grad_mask beyond the pixel's own group."""
    c = cases.Case("inf_grad", N=1, H=40, W=40, G=8, gc=16, seed=93)
    v, o, m, g = cases.make_inputs(c)
    py, px, ch = 20, 17, 3 * 16 + 5
    g[0, py, px, ch] = np.inf
    arrs = rounded((v, o / 3.0, m, g), torch.bfloat16)
    got = run_cuda(arrs, c.geom, dtype=torch.bfloat16)
    gv, go, gm = got[1], got[2], got[3]
    bad = ~np.isfinite(gv)
    assert bad.any()
    n_, ys, xs, cs = np.nonzero(bad)
    assert set(cs.tolist()) == {ch}
    assert ys.min() >= py - 16 and ys.max() <= py + 16 and xs.min() >= px - 16 and xs.max() <= px + 16
    bad_o = ~np.isfinite(go.reshape(1, c.H, c.W, c.G, 18))
    bad_m = ~np.isfinite(gm.reshape(1, c.H, c.W, c.G, 9))
    for b in (bad_o, bad_m):
        idx = np.argwhere(b)
        assert len(idx) and all((i[1], i[2], i[3]) == (py, px, 3) for i in idx)
    # everything that stayed finite is still right
    want = _oracle(tuple(np.where(np.isfinite(a), a, 0.0) for a in arrs), c.geom)
    ok = ~bad
    rms = float(np.sqrt(np.mean(want[1] ** 2)))
    far = np.ones_like(ok); far[:, py - 16:py + 17, px - 16:px + 17, ch] = False
    assert allclose_frac(gv[ok & far], want[1][ok & far], rtol=1e-2, atol=1e-2 * rms) <= 1e-3


# ----------------------------------------------------------------------------- the reference's own defaults
@pytest.mark.parametrize("dt", ["bf16", "f16", "f32"])
@pytest.mark.parametrize("name,kw", [("layer_default_c64_g4", dict(G=4, gc=16)),
                                     ("reftest_sigma2", dict(G=8, gc=16, sigma=2.0)),
                                     ("stride2", dict(G=8, gc=16, sh=2, sw=2))])
def test_reference_default_shapes(name, kw, dt):
    """Shapes the reference itself uses and that fall outside the fastest kernels' gates: the layer's constructor
    default channels=64, group=4 (modules/dcnv3.py:223-237), the test script's offset_scale=2.0 (test.py:19-30), and
    a stride-2 layer.  Parity only (their timing rows are in profiles/README.md)."""
    c = cases.Case(name, N=2, H=30, W=28, seed=94, **kw)
    arrs = rounded(cases.make_inputs(c), TDT[dt])
    got = run_cuda(arrs, c.geom, dtype=TDT[dt])
    want = _oracle(arrs, c.geom)
    if dt == "f32":
        for nm, a, w in zip(WHAT, got, want):
            lim = 1e-3 if nm == "go" else 0.0
            assert allclose_frac(a, w, rtol=1e-5, atol=1e-5 * (1.0 + float(np.sqrt(np.mean(w ** 2))))) <= lim, (nm, max_abs(a, w))
    else:
        _check_16bit(got, want, dt)


# ----------------------------------------------------------------------------- fused producer backward kernels
@pytest.mark.parametrize("dt", ["bf16", "f16"])
@pytest.mark.parametrize("N,H,W,C", [(2, 16, 24, 256), (1, 13, 19, 128), (3, 9, 8, 64)])
def test_fused_dwconv_ln_gelu_backward_matches_fp64_autograd(N, H, W, C, dt):
    """csrc/dcnv3_dwconv_bwd.cu (LayerNorm + GELU backward with the channel sums, depthwise dgrad + wgrad on TMA windows)
    against autograd of conv2d(groups=C) -> layer_norm -> gelu in fp64 on the rounded inputs (modules/dcnv3.py:276-289):
    all five gradients, 1e-2 relative with a floor of 1e-2 x RMS; ragged tiles included."""
    from yolo_somi_b200.ops_dcnv3.functions import dwconv_ln_gelu as dlg
    dtype = TDT[dt]
    g = torch.Generator(device="cpu").manual_seed(N * H + C)
    x = torch.randn(N, H, W, C, generator=g).to(dtype)
    w = (torch.randn(C, 1, 3, 3, generator=g) / 3).to(dtype)
    b, gamma, beta = torch.randn(C, generator=g), 1 + 0.3 * torch.randn(C, generator=g), torch.randn(C, generator=g)
    go = torch.randn(N, H, W, C, generator=g).to(dtype)
    args = [t.cuda().requires_grad_(True) for t in (x, w.float(), b, gamma, beta)]
    out = dlg.DwConvLnGelu.apply(*args, 1e-6, dtype)
    out.backward(go.cuda())
    torch.cuda.synchronize()
    ref_in = [t.double().requires_grad_(True) for t in (x, w, b, gamma, beta)]
    # the kernel's backward starts from the convolution output as the forward stored it (rounded to the I/O dtype)
    want = torch.autograd.grad(dlg._unfused(*ref_in, 1e-6), ref_in, go.double())
    for name, a, wv in zip(("grad_x", "grad_w", "grad_b", "grad_gamma", "grad_beta"), args, want):
        a, wv = a.grad.double().cpu().numpy(), wv.numpy()
        rms = float(np.sqrt(np.mean(wv ** 2)))
        # reductions over N H W pixels of 16-bit-rounded du: relative L2 for the parameter gradients
        if name == "grad_x":
            assert allclose_frac(a, wv, rtol=2e-2, atol=2e-2 * rms) <= 1e-3, (name, max_abs(a, wv), rms)
        else:
            assert float(np.linalg.norm(a - wv) / (np.linalg.norm(wv) + 1e-30)) <= 1e-2, name


@pytest.mark.parametrize("rows", [4321, 4320])      # rows % 8 == 0: the 128-bit form (eight rows per thread)
@pytest.mark.parametrize("dt", ["bf16", "f16"])
def test_mask_softmax_backward_kernel(dt, rows):
    from yolo_somi_b200 import _native
    lib = _native.load()
    dtype = TDT[dt]
    P = 9
    gen = torch.Generator(device="cpu").manual_seed(5)
    mask = torch.softmax(torch.randn(rows, P, generator=gen), -1).to(dtype)
    gm = torch.randn(rows, P, generator=gen).to(dtype)
    dm, dg = mask.cuda(), gm.cuda()
    out = torch.empty_like(dm)
    rc = lib.dcnv3_mask_softmax_backward_sm100(dg.data_ptr(), dm.data_ptr(), out.data_ptr(), rows, P,
                                               _native.BF16 if dt == "bf16" else _native.F16,
                                               torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    want = torch._softmax_backward_data(gm.double(), mask.double(), -1, torch.float64)
    eps = 2.0 ** -8 if dt == "bf16" else 2.0 ** -11
    assert float((out.double().cpu() - want).abs().max()) <= 1.01 * eps * float(want.abs().max()) + 1e-9
    if rows % 8 == 0:
        # bit-identical to the scalar form: the same rows inside a call whose row count is not a multiple of eight
        dm2, dg2 = torch.cat([dm, dm[:1]]), torch.cat([dg, dg[:1]])
        out2 = torch.empty_like(dm2)
        rc = lib.dcnv3_mask_softmax_backward_sm100(dg2.data_ptr(), dm2.data_ptr(), out2.data_ptr(), rows + 1, P,
                                                   _native.BF16 if dt == "bf16" else _native.F16,
                                                   torch.cuda.current_stream().cuda_stream)
        assert rc == 0 and torch.equal(out2[:rows].view(torch.int16), out.view(torch.int16))


# ----------------------------------------------------------------------------- the far-point list, mid range
@pytest.mark.gpu
@pytest.mark.parametrize("shape,spread,dt", [((2, 80, 80, 8), 1.8, torch.bfloat16), ((8, 80, 80, 16), 2.0, torch.bfloat16),
                                              ((4, 64, 96, 8), 2.0, torch.float16)],
                         ids=["14k-far-bf16", "170k-far-bf16", "50k-far-f16"])
def test_far_point_list_between_a_few_and_the_fallback_threshold(shape, spread, dt):
    """Offsets a couple of pixels wide put 1 - 3 % of the points beyond their patch's band: below the 1 / 32 at which the
    plane form takes over, far above the few thousand of N(0, 1) offsets -- `far_points` then handles several entries per
    warp and per batch (its lane <-> entry phase), which no other test reaches.  Against the C oracle in fp64 on the
    rounded inputs, the bars of the other 16-bit tests."""
    from oracle import dcnv3_oracle as orc
    from helpers import allclose_frac, max_abs
    n, h, w, G = shape
    gen = torch.Generator(device="cpu").manual_seed(31 + n)
    value = torch.randn(n, h, w, G * 16, generator=gen)
    offset = spread * torch.randn(n, h, w, G * 18, generator=gen)
    mask = torch.softmax(torch.randn(n, h, w, G, 9, generator=gen), -1).reshape(n, h, w, G * 9)
    grad = torch.randn(n, h, w, G * 16, generator=gen)
    geom = (3, 3, 1, 1, 1, 1, 1, 1, G, 16, 1.0)
    dev = [t.to(dt).cuda() for t in (value, offset, mask, grad)]
    arrs = [t.double().cpu().numpy() for t in dev]
    # the share of points beyond the band (dcnv3_backward_vres.cu: 0 <= px + i + 3 + dx < 15 on both axes) must sit in the
    # range this test is about
    o = arrs[1].reshape(n, h, w, G, 9, 2)
    px = (np.arange(w) % 8)[None, None, :, None, None]
    py = (np.arange(h) % 8)[None, :, None, None, None]
    pi = (np.arange(9) // 3)[None, None, None, None, :]
    pj = (np.arange(9) % 3)[None, None, None, None, :]
    ub, vb = px + pi + 3 + o[..., 0], py + pj + 3 + o[..., 1]
    far = 1.0 - np.mean((ub >= 0) & (ub < 15) & (vb >= 0) & (vb < 15))
    assert 5e-3 < far < 1.0 / 32, far
    import DCNv3
    got = list(DCNv3.dcnv3_backward(*dev[:3], *geom, dev[3], 256))
    torch.cuda.synchronize()
    want = orc.direct_backward(*arrs, *geom)
    for name, a, wnt in zip(("gv", "go", "gm"), got, want):
        a = a.double().cpu().numpy()
        rms = float(np.sqrt(np.mean(wnt ** 2)))
        frac = allclose_frac(a, wnt, rtol=1e-2, atol=1e-2 * rms)
        assert frac <= (5e-4 if name == "go" else 2e-4), (name, frac, max_abs(a, wnt), rms)
        if name != "go":
            worst = float(np.max(np.abs(a - wnt) / np.maximum(np.abs(wnt), rms)))
            assert worst <= (8e-2 if dt == torch.bfloat16 else 2e-2), (name, worst)


@pytest.mark.gpu
def test_small_shapes_take_the_plane_form_by_default():
    """The size heuristic of the default dispatch (backward_vres_preferred: >= 8 patch rows per SM, >= 4 patches per row)
    in a FRESH process without the test suite's DCNV3_VRES_MIN_ROWS=0: a small shape and a large one, both against the
    C oracle -- whichever form runs, the result is the same function."""
    import subprocess, sys, textwrap
    code = textwrap.dedent('''
        import sys, numpy as np, torch
        sys.path[:0] = [%r, %r]
        import DCNv3
        from oracle import dcnv3_oracle as orc
        for n, h, w, G in ((2, 40, 40, 8), (12, 80, 80, 16)):
            g = torch.Generator().manual_seed(5)
            v = torch.randn(n, h, w, G * 16, generator=g); o = torch.randn(n, h, w, G * 18, generator=g)
            m = torch.softmax(torch.randn(n, h, w, G, 9, generator=g), -1).reshape(n, h, w, -1); go = torch.randn(n, h, w, G * 16, generator=g)
            dev = [t.bfloat16().cuda() for t in (v, o, m, go)]
            geom = (3, 3, 1, 1, 1, 1, 1, 1, G, 16, 1.0)
            got = DCNv3.dcnv3_backward(*dev[:3], *geom, dev[3], 256)
            arrs = [t.double().cpu().numpy() for t in dev]
            want = orc.direct_backward(*arrs, *geom)
            for a, wnt in zip(got, want):
                a = a.double().cpu().numpy(); rms = float(np.sqrt(np.mean(wnt ** 2)))
                assert np.mean(np.abs(a - wnt) > 1e-2 * np.abs(wnt) + 1e-2 * rms) <= 5e-4
        print("ok")
    ''') % (str(cases.__file__).rsplit("/tests/", 1)[0], str(cases.__file__).rsplit("/golden/", 1)[0])
    env = {k: v for k, v in __import__("os").environ.items() if k != "DCNV3_VRES_MIN_ROWS"}
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
