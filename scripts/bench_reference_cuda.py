#!/usr/bin/env python
"""The reference's own CUDA kernels, recompiled for sm_100a, timed beside this library on the same box.

    python scripts/build_reference_cuda.py      # once, in the build container (-> baseline/_ref/DCNv3_refcuda.so)
    python scripts/bench_reference_cuda.py      # on a GPU box; prints one JSON object

`DCNv3_refcuda` is the reference's `models/ops_dcnv3/src` with the two `input.type()` tokens of its AT_DISPATCH calls
replaced (the extension does not compile against torch 2.11 otherwise) and nothing else changed: its im2col forward
(`dcnv3_im2col_cuda.cuh:216-275`), its col2im backward with global fp32 atomics (`:82-147,278-370`) and its host
launchers with their zero fills and cast passes (`dcnv3_cuda.cu:55-57,126-133,168-173`).  It has no bf16 dispatch, so
the comparison runs in fp16 and fp32 at BASELINE configs[1]'s shape, and in fp32 at configs[0]'s.  Beside the timings
the two implementations' results are compared element by element (the reference's own test tolerances,
`test.py:85,134-148`: rtol 1e-2 / atol 1e-3) -- parity against the reference's GPU path, not only its CPU oracle.
"""
from __future__ import annotations

import importlib.util
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
SO = ROOT / "baseline" / "_ref" / "DCNv3_refcuda.so"


def load_reference_ext():
    if not SO.exists():
        return None
    spec = importlib.util.spec_from_file_location("DCNv3_refcuda", SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_inputs(n, h, w, c, g, dtype, dev, seed, count=3):
    sets = []
    for i in range(count):
        gen = torch.Generator(device="cpu").manual_seed(seed + i)
        value = torch.randn(n, h, w, c, generator=gen)
        offset = torch.randn(n, h, w, g * 18, generator=gen)
        mask = torch.softmax(torch.randn(n, h, w, g, 9, generator=gen), -1).reshape(n, h, w, g * 9)
        grad = torch.randn(n, h, w, c, generator=gen)
        sets.append(tuple(t.to(dtype).to(dev) for t in (value, offset, mask, grad)))
    return sets


def time_impl(mod, sets, geo, iters):
    """CUDA events around `iters` forward and `iters` backward calls over rotating input sets; ms per call."""
    for v, o, m, go in sets[:2]:
        mod.dcnv3_forward(v, o, m, *geo, 256)
        mod.dcnv3_backward(v, o, m, *geo, go, 256)
    torch.cuda.synchronize()
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    for i in range(iters):
        v, o, m, go = sets[i % len(sets)]
        mod.dcnv3_forward(v, o, m, *geo, 256)
    b.record()
    for i in range(iters):
        v, o, m, go = sets[i % len(sets)]
        mod.dcnv3_backward(v, o, m, *geo, go, 256)
    c.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, b.elapsed_time(c) / iters


def compare(ref, ours, sets, geo):
    """Element-wise agreement of the two GPU implementations: violation fraction of |a - b| <= atol + rtol |b| at the
    reference test's tolerances, and the largest difference relative to the tensor's RMS."""
    v, o, m, go = sets[0]
    r = [ref.dcnv3_forward(v, o, m, *geo, 256)] + list(ref.dcnv3_backward(v, o, m, *geo, go, 256))
    u = [ours.dcnv3_forward(v, o, m, *geo, 256)] + list(ours.dcnv3_backward(v, o, m, *geo, go, 256))
    out = {}
    for name, x, y in zip(("out", "grad_value", "grad_offset", "grad_mask"), u, r):
        x, y = x.double(), y.double()
        d = (x - y).abs()
        out[name] = {"violations_rtol1e-2_atol1e-3": float((d > 1e-3 + 1e-2 * y.abs()).double().mean()),
                     "max_abs_over_rms": float(d.max() / y.pow(2).mean().sqrt())}
    return out


def run(iters=10, dev=None):
    ref = load_reference_ext()
    if ref is None:
        return {"unavailable": "baseline/_ref/DCNv3_refcuda.so not built (scripts/build_reference_cuda.py)"}
    import DCNv3 as ours
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    rows = {}
    cases = [("cfg2_fp16", (16, 80, 80, 256, 16), torch.float16), ("cfg2_fp32", (16, 80, 80, 256, 16), torch.float32),
             ("cfg1_fp32", (2, 40, 40, 64, 4), torch.float32)]
    for name, (n, h, w, c, g), dtype in cases:
        geo = (3, 3, 1, 1, 1, 1, 1, 1, g, c // g, 1.0)
        sets = make_inputs(n, h, w, c, g, dtype, dev, seed=7)
        rf, rb = time_impl(ref, sets, geo, iters)
        of, ob = time_impl(ours, sets, geo, iters)
        pts = n * h * w * g * 9
        rows[name] = {"reference_cuda": {"fwd_ms": rf, "bwd_ms": rb, "pts_per_s": pts / ((rf + rb) * 1e-3)},
                      "this_library": {"fwd_ms": of, "bwd_ms": ob, "pts_per_s": pts / ((of + ob) * 1e-3)},
                      "speedup_fwd_bwd": (rf + rb) / (of + ob), "agreement": compare(ref, ours, sets, geo)}
        del sets
        torch.cuda.empty_cache()
    return {"what": "the reference's ops_dcnv3 CUDA extension recompiled for sm_100a (two-token torch-2.11 patch, "
                    "scripts/build_reference_cuda.py) against this library, same box, same inputs, CUDA events, "
                    "%d calls per pass over 3 rotating input sets; the reference has no bf16 dispatch" % iters,
            "rows": rows}


if __name__ == "__main__":
    if not torch.cuda.is_available():
        raise SystemExit("needs a GPU")
    print(json.dumps(run()))
