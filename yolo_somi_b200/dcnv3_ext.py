"""Tensor-level entry points with the reference extension's exact signatures.

The reference builds a pybind11 module called ``DCNv3`` that exports two functions
(models/ops_dcnv3/src/vision.cpp:14-17, models/ops_dcnv3/src/dcnv3.h:20-59) and calls them from
``DCNv3Function`` (models/ops_dcnv3/functions/dcnv3_func.py:39-43,53-58).  This module provides
the same two callables -- same positional arguments, same return types, same error conditions
(``RuntimeError``) -- on top of the C ABI of libdcnv3_sm100.so.  The repo-root ``DCNv3.py``
re-exports them so that ``import DCNv3`` resolves here and the reference's Python files run
unmodified.

Deviations from the reference, all documented in DESIGN.md:
  * bf16 is supported beside the reference's fp64/fp32/fp16 (dcnv3_cuda.cu:69); fp64 is a plain all-double
    correctness path (csrc/dcnv3_f64.cu: the reference's test.py drives the extension in double) without a
    deterministic mode;
  * a failed kernel launch raises (the reference printf()s, dcnv3_im2col_cuda.cuh:864-867);
  * outputs are allocated with ``torch.empty`` (every element is written by the kernels);
  * half-precision gradients are produced directly (no fp32 round trip of grad_offset/grad_mask);
  * ``im2col_step`` only takes part in the divisibility check: one launch covers the batch;
  * ``DCNV3_DETERMINISTIC=1`` (or torch.use_deterministic_algorithms(True)) selects the
    bit-reproducible backward.
"""
from __future__ import annotations

import os

import torch

from . import _native

_DTYPES = {torch.float32: _native.F32, torch.float16: _native.F16, torch.bfloat16: _native.BF16, torch.float64: _native.F64}


def _conv_out(size, pad, dil, k, stride):
    return (size + 2 * pad - (dil * (k - 1) + 1)) // stride + 1  # dcnv3_cuda.cu:40-45


def _validate(tensors, group, group_channels, im2col_step):
    """Same conditions as dcnv3_cuda.cu:29-53 / dcnv3.h:37."""
    inp = tensors[0][1]
    for name, t in tensors:
        if not isinstance(t, torch.Tensor):
            raise TypeError(f"{name} must be a torch.Tensor")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
    if not inp.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    for name, t in tensors:
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")
        if t.device != inp.device or t.dtype != inp.dtype:
            raise RuntimeError(f"{name}: expected {inp.dtype} on {inp.device}, got {t.dtype} on {t.device}")
    if inp.dtype not in _DTYPES:
        raise RuntimeError(f"dcnv3: unsupported dtype {inp.dtype} (float32, float16, bfloat16, float64)")
    if inp.dim() != 4:
        raise RuntimeError("input must be [N, H, W, C]")
    batch, channels = inp.shape[0], inp.shape[3]
    step = min(batch, int(im2col_step))
    if batch > 0 and (step <= 0 or batch % step != 0):
        raise RuntimeError(f"batch({batch}) must divide im2col_step({step})")
    if channels != group * group_channels:
        raise RuntimeError("Input channels and group times group channels wont match: "
                           f"({channels} vs {group * group_channels}).")


def _geometry(inp, offset, mask, kh, kw, sh, sw, ph, pw, dh, dw, group, gc):
    n, h, w, _ = inp.shape
    ho, wo = _conv_out(h, ph, dh, kh, sh), _conv_out(w, pw, dw, kw, sw)
    pts = group * kh * kw
    if tuple(offset.shape) != (n, ho, wo, pts * 2):
        raise RuntimeError(f"offset must be {(n, ho, wo, pts * 2)}, got {tuple(offset.shape)}")
    if tuple(mask.shape) != (n, ho, wo, pts):
        raise RuntimeError(f"mask must be {(n, ho, wo, pts)}, got {tuple(mask.shape)}")
    return (n, h, w, ho, wo, group, gc, kh, kw, sh, sw, ph, pw, dh, dw)


def _aligned(t: torch.Tensor) -> torch.Tensor:
    # sliced views can start off a 16-byte boundary; the kernels want aligned bases
    return t if t.data_ptr() % 16 == 0 else t.clone(memory_format=torch.contiguous_format)


class _OnDevice:
    """`with torch.cuda.device(dev)` only when `dev` is not already current: the context manager costs ~5 us per call,
    a visible share of a launch-bound call on a 20 x 20 feature map (scripts/small_map.py)."""
    __slots__ = ("ctx",)

    def __init__(self, device):
        self.ctx = None if device.index is None or device.index == torch.cuda.current_device() else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)


def deterministic_requested() -> bool:
    return os.environ.get("DCNV3_DETERMINISTIC", "0") not in ("", "0") or \
        torch.are_deterministic_algorithms_enabled()


def dcnv3_forward(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                  dilation_h, dilation_w, group, group_channels, offset_scale, im2col_step):
    """Drop-in for ``DCNv3.dcnv3_forward`` (src/dcnv3.h:20-38): returns out [N,Ho,Wo,C]."""
    _validate((("input", input), ("offset", offset), ("mask", mask)), group, group_channels,
              im2col_step)
    geom = _geometry(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                     dilation_h, dilation_w, group, group_channels)
    lib = _native.load()
    input, offset, mask = _aligned(input), _aligned(offset), _aligned(mask)
    out = torch.empty((geom[0], geom[3], geom[4], group * group_channels),
                      dtype=input.dtype, device=input.device)
    with _OnDevice(input.device):
        stream = torch.cuda.current_stream().cuda_stream
        rc = lib.dcnv3_forward_sm100(input.data_ptr(), offset.data_ptr(), mask.data_ptr(),
                                     out.data_ptr(), *geom, float(offset_scale),
                                     _DTYPES[input.dtype], stream)
    _native.check(rc, "dcnv3_forward_sm100")
    return out


def dcnv3_backward(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                   dilation_h, dilation_w, group, group_channels, offset_scale, grad_output,
                   im2col_step):
    """Drop-in for ``DCNv3.dcnv3_backward`` (src/dcnv3.h:40-59):
    returns [grad_input, grad_offset, grad_mask] in the dtype of the inputs."""
    _validate((("input", input), ("offset", offset), ("mask", mask),
               ("grad_output", grad_output)), group, group_channels, im2col_step)
    geom = _geometry(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                     dilation_h, dilation_w, group, group_channels)
    if tuple(grad_output.shape) != (geom[0], geom[3], geom[4], group * group_channels):
        raise RuntimeError(f"grad_output has shape {tuple(grad_output.shape)}")
    lib = _native.load()
    input, offset, mask, grad_output = map(_aligned, (input, offset, mask, grad_output))
    grad_input = torch.empty_like(input)
    grad_offset = torch.empty_like(offset)
    grad_mask = torch.empty_like(mask)
    flags = _native.BWD_DETERMINISTIC if deterministic_requested() else 0
    dt = _DTYPES[input.dtype]
    with _OnDevice(input.device):
        nbytes = lib.dcnv3_backward_workspace_bytes(geom[0], geom[1], geom[2], group,
                                                    group_channels, dt, flags)
        work = torch.empty(nbytes, dtype=torch.uint8, device=input.device) if nbytes else None
        stream = torch.cuda.current_stream().cuda_stream
        rc = lib.dcnv3_backward_sm100(input.data_ptr(), offset.data_ptr(), mask.data_ptr(),
                                      grad_output.data_ptr(), grad_input.data_ptr(),
                                      grad_offset.data_ptr(), grad_mask.data_ptr(),
                                      work.data_ptr() if work is not None else None, nbytes,
                                      *geom, float(offset_scale), dt, flags, stream)
    _native.check(rc, "dcnv3_backward_sm100")
    return [grad_input, grad_offset, grad_mask]
