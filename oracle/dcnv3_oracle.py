"""oracle/dcnv3_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU oracle for the DCNv3 core (grouped deformable bilinear sampling x softmax
modulation mask).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module; nothing under ``yolo_somi_b200/`` does, and the product path raises if
its CUDA library is missing rather than falling back to anything here.

Two restatements of the reference algorithm (paths relative to /root/reference):

``core_gridsample``
    the reference's *semantics of record*: ``dcnv3_core_pytorch``,
    models/ops_dcnv3/functions/dcnv3_func.py:147-188, with its helpers
    ``_get_reference_points`` (:91-119) and ``_generate_dilation_grids``
    (:122-144).  The arithmetic lives in ``torch.nn.functional.grid_sample``
    (third-party: PyTorch; 2.11.0 in this image, the reference pinned 1.13.1 in
    requirements.txt:147) with ``mode='bilinear', padding_mode='zeros',
    align_corners=False``.  Gradients come from autograd, as they do for the
    reference's own test (models/ops_dcnv3/test.py:93-216).  This is also the
    CPU baseline that ``bench.py`` times ("port" of the reference CPU path).

``direct_forward`` / ``direct_backward``
    the pixel-space formulas of the reference CUDA extension
    (models/ops_dcnv3/src/cuda/dcnv3_im2col_cuda.cuh:32-147,216-275), in C
    (oracle/dcnv3_direct.c), fp64 "truth" and fp32.

Parity pin: both are checked against golden vectors produced by importing the
reference's own Python implementation in the build container
(tests/golden/make_golden.py -> tests/golden/*.npz; tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libdcnv3_oracle.so"
_lib = None


# --------------------------------------------------------------------------- geometry
def out_size(size_in: int, pad: int, dil: int, k: int, stride: int) -> int:
    """models/ops_dcnv3/src/cuda/dcnv3_cuda.cu:40-45."""
    return (size_in + 2 * pad - (dil * (k - 1) + 1)) // stride + 1


# --------------------------------------------------------------------------- grid_sample form
def _anchor_axis(n_out: int, k: int, dil: int, stride: int, extent: int) -> torch.Tensor:
    """Normalised centre of every output position along one axis of the padded map.

    dcnv3_func.py:96-115 -- fp32 linspace from (dil*(k-1))//2 + 0.5 in steps of
    ``stride``, divided by the padded extent.
    """
    first = (dil * (k - 1)) // 2 + 0.5
    pts = torch.linspace(first, first + (n_out - 1) * stride, n_out, dtype=torch.float32)
    return pts / extent


def _kernel_axis(k: int, dil: int, extent: int) -> torch.Tensor:
    """Normalised kernel tap positions along one axis (dcnv3_func.py:125-139)."""
    first = -((dil * (k - 1)) // 2)
    pts = torch.linspace(first, first + (k - 1) * dil, k, dtype=torch.float32)
    return pts / extent


def core_gridsample(value, offset, mask, kernel_h, kernel_w, stride_h, stride_w,
                    pad_h, pad_w, dilation_h, dilation_w, group, group_channels,
                    offset_scale):
    """Restatement of ``dcnv3_core_pytorch`` (dcnv3_func.py:147-188).  Differentiable.

    value [N,H,W,G*gc]; offset [N,Ho,Wo,G*P*2] as (dx,dy) per point; mask
    [N,Ho,Wo,G*P]; returns [N,Ho,Wo,G*gc].
    """
    # dcnv3_func.py:154-156 pads the W axis by pad_h and the H axis by pad_w
    # (F.pad lists the last dimension first); identical when pad_h == pad_w.
    padded = F.pad(value, [0, 0, pad_h, pad_h, pad_w, pad_w])
    n, hp, wp, _ = padded.shape
    ho, wo = offset.shape[1], offset.shape[2]
    pts = kernel_h * kernel_w
    # the reference derives its own output size from the padded map (:93-94)
    ho_ref = (hp - (dilation_h * (kernel_h - 1) + 1)) // stride_h + 1
    wo_ref = (wp - (dilation_w * (kernel_w - 1) + 1)) // stride_w + 1

    ay = _anchor_axis(ho_ref, kernel_h, dilation_h, stride_h, hp)          # [Ho]
    ax = _anchor_axis(wo_ref, kernel_w, dilation_w, stride_w, wp)          # [Wo]
    anchor = torch.stack((ax[None, :].expand(ho_ref, wo_ref),
                          ay[:, None].expand(ho_ref, wo_ref)), -1)         # [Ho,Wo,(x,y)]
    anchor = anchor.reshape(1, ho_ref, wo_ref, 1, 2)

    kx = _kernel_axis(kernel_w, dilation_w, wp)                            # [Kw]
    ky = _kernel_axis(kernel_h, dilation_h, hp)                            # [Kh]
    # point p = i*Kh + j, i over width (outer), j over height (inner)  (:124-139)
    taps = torch.stack((kx[:, None].expand(kernel_w, kernel_h),
                        ky[None, :].expand(kernel_w, kernel_h)), -1).reshape(pts, 2)
    taps = taps[None].expand(group, pts, 2).reshape(1, 1, 1, group * pts, 2)

    extent = torch.tensor([wp, hp]).reshape(1, 1, 1, 2).repeat(1, 1, 1, group * pts)

    loc = (anchor + taps * offset_scale).repeat(n, 1, 1, 1, 1).flatten(3, 4) \
        + offset * offset_scale / extent                                   # (:167-168)
    grid = 2 * loc - 1

    planes = padded.view(n, hp * wp, group * group_channels).transpose(1, 2) \
        .reshape(n * group, group_channels, hp, wp)
    grid = grid.view(n, ho * wo, group, pts, 2).transpose(1, 2).flatten(0, 1)
    sampled = F.grid_sample(planes, grid, mode="bilinear", padding_mode="zeros",
                            align_corners=False)                           # [N*G,gc,Ho*Wo,P]
    weights = mask.view(n, ho * wo, group, pts).transpose(1, 2) \
        .reshape(n * group, 1, ho * wo, pts)
    mixed = (sampled * weights).sum(-1).view(n, group * group_channels, ho * wo)
    return mixed.transpose(1, 2).reshape(n, ho, wo, -1).contiguous()


def gridsample_fwd_bwd(value, offset, mask, grad_out, *geom):
    """Forward + autograd backward of ``core_gridsample``; returns (out, gv, go, gm)."""
    v = value.detach().clone().requires_grad_(True)
    o = offset.detach().clone().requires_grad_(True)
    m = mask.detach().clone().requires_grad_(True)
    out = core_gridsample(v, o, m, *geom)
    out.backward(grad_out)
    return out.detach(), v.grad, o.grad, m.grad


# --------------------------------------------------------------------------- direct (C) form
class _Geom(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int) for k in
                ("N", "H", "W", "Ho", "Wo", "G", "gc", "kh", "kw", "sh", "sw",
                 "ph", "pw", "dh", "dw")] + [("offset_scale", ctypes.c_double)]


def build(force: bool = False) -> Path:
    """Compile oracle/dcnv3_direct.c -> oracle/libdcnv3_oracle.so (gcc, OpenMP)."""
    srcs = [_HERE / "dcnv3_direct.c", _HERE / "dcnv3_direct_impl.inc"]
    if (not force and _LIB_PATH.exists()
            and all(_LIB_PATH.stat().st_mtime >= s.stat().st_mtime for s in srcs if s.exists())):
        return _LIB_PATH
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
           "-o", str(_LIB_PATH), str(srcs[0]), "-lm"]
    subprocess.run(cmd, check=True, cwd=str(_HERE))
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            build()
        _lib = ctypes.CDLL(str(_LIB_PATH))
        for sfx in ("f64", "f32"):
            getattr(_lib, f"dcnv3_oracle_forward_{sfx}").restype = ctypes.c_int
            getattr(_lib, f"dcnv3_oracle_backward_{sfx}").restype = ctypes.c_int
    return _lib


def _geom(value, offset, kh, kw, sh, sw, ph, pw, dh, dw, group, gc, sigma) -> _Geom:
    n, h, w, c = value.shape
    assert c == group * gc, (c, group, gc)
    ho, wo = out_size(h, ph, dh, kh, sh), out_size(w, pw, dw, kw, sw)
    assert tuple(offset.shape) == (n, ho, wo, group * kh * kw * 2), (offset.shape, ho, wo)
    return _Geom(n, h, w, ho, wo, group, gc, kh, kw, sh, sw, ph, pw, dh, dw, float(sigma))


def _np(t, dtype):
    a = t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def direct_forward(value, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, group, gc, sigma,
                   dtype=np.float64) -> np.ndarray:
    """Pixel-space forward (dcnv3_im2col_cuda.cuh:216-275) in ``dtype`` arithmetic."""
    lib = _load()
    sfx = "f64" if dtype == np.float64 else "f32"
    v, o, m = _np(value, dtype), _np(offset, dtype), _np(mask, dtype)
    q = _geom(v, o, kh, kw, sh, sw, ph, pw, dh, dw, group, gc, sigma)
    out = np.empty((q.N, q.Ho, q.Wo, group * gc), dtype=dtype)
    rc = getattr(lib, f"dcnv3_oracle_forward_{sfx}")(_ptr(v), _ptr(o), _ptr(m), _ptr(out),
                                                      ctypes.byref(q))
    assert rc == 0
    return out


def direct_backward(value, offset, mask, grad_out, kh, kw, sh, sw, ph, pw, dh, dw, group, gc,
                    sigma, dtype=np.float64):
    """Pixel-space gradients (dcnv3_im2col_cuda.cuh:82-147): (g_value, g_offset, g_mask)."""
    lib = _load()
    sfx = "f64" if dtype == np.float64 else "f32"
    v, o, m, g = (_np(t, dtype) for t in (value, offset, mask, grad_out))
    q = _geom(v, o, kh, kw, sh, sw, ph, pw, dh, dw, group, gc, sigma)
    assert g.shape == (q.N, q.Ho, q.Wo, group * gc)
    gv, go, gm = np.empty_like(v), np.empty_like(o), np.empty_like(m)
    rc = getattr(lib, f"dcnv3_oracle_backward_{sfx}")(_ptr(v), _ptr(o), _ptr(m), _ptr(g),
                                                       _ptr(gv), _ptr(go), _ptr(gm),
                                                       ctypes.byref(q))
    assert rc == 0
    return gv, go, gm


if __name__ == "__main__":
    print(build(force=True))
