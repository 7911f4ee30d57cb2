"""Host-buffer entry point of the DCNv3 core: forward + backward with every tensor in (pinned)
host memory, chunked over the batch and pipelined over three CUDA streams inside
libdcnv3_sm100.so (``dcnv3_host_pipeline_*`` in include/dcnv3_sm100.h).

The reference moves a step's tensors with ``imgs.to(device, non_blocking=True)`` (train.py:249),
calls the op (models/ops_dcnv3/functions/dcnv3_func.py:39-43,53-58) and reads results back, one
after the other on one stream.  Here the H2D copy of chunk c+1, the kernels of chunk c and the D2H
copy of chunk c-1 overlap, also across consecutive ``run`` calls; ``sync`` waits for everything.

    pipe = DCNv3HostPipeline(H=80, W=80, group=16, group_channels=16, dtype=torch.bfloat16)
    pipe.run(value, offset, mask, grad_out, out, grad_value, grad_offset, grad_mask)   # CPU tensors
    pipe.sync()
"""
from __future__ import annotations

import ctypes

import torch

from . import _native
from .dcnv3_ext import _DTYPES, _conv_out


class DCNv3HostPipeline:
    def __init__(self, H, W, group, group_channels, kernel=3, stride=1, pad=1, dilation=1,
                 offset_scale=1.0, dtype=torch.bfloat16, chunk_images=4, deterministic=False,
                 device=None):
        if dtype not in _DTYPES:
            raise RuntimeError(f"dcnv3: unsupported dtype {dtype} (float32, float16, bfloat16)")
        self._lib = _native.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dtype, self.chunk = dtype, int(chunk_images)
        self.H, self.W, self.G, self.gc, self.K = H, W, group, group_channels, kernel
        self.Ho, self.Wo = _conv_out(H, pad, dilation, kernel, stride), _conv_out(W, pad, dilation, kernel, stride)
        self._h = ctypes.c_void_p()
        flags = _native.BWD_DETERMINISTIC if deterministic else 0
        with torch.cuda.device(self.device):
            rc = self._lib.dcnv3_host_pipeline_create(
                ctypes.byref(self._h), self.chunk, H, W, group, group_channels, kernel, kernel,
                stride, stride, pad, pad, dilation, dilation, float(offset_scale), _DTYPES[dtype], flags)
        _native.check(rc, "dcnv3_host_pipeline_create")

    def shapes(self, n):
        """(value/grad_value, offset/grad_offset, mask/grad_mask, out/grad_out) shapes for n images."""
        C, P = self.G * self.gc, self.K * self.K
        return ((n, self.H, self.W, C), (n, self.Ho, self.Wo, self.G * P * 2),
                (n, self.Ho, self.Wo, self.G * P), (n, self.Ho, self.Wo, C))

    def run(self, value, offset, mask, grad_out, out, grad_value, grad_offset, grad_mask):
        """Enqueue forward + backward of the whole batch; CPU tensors (ideally pinned) of this
        pipeline's dtype and the shapes of ``shapes(N)``.  Returns at once; call ``sync()``."""
        n = value.shape[0]
        sv, so, sm, sy = self.shapes(n)
        for name, t, shp in (("value", value, sv), ("offset", offset, so), ("mask", mask, sm),
                             ("grad_out", grad_out, sy), ("out", out, sy), ("grad_value", grad_value, sv),
                             ("grad_offset", grad_offset, so), ("grad_mask", grad_mask, sm)):
            if t.is_cuda or t.dtype != self.dtype or tuple(t.shape) != shp or not t.is_contiguous():
                raise RuntimeError(f"{name}: expected a contiguous CPU {self.dtype} tensor of shape {shp}, "
                                   f"got {t.dtype} {tuple(t.shape)} on {t.device}")
        with torch.cuda.device(self.device):
            rc = self._lib.dcnv3_host_pipeline_run(
                self._h, value.data_ptr(), offset.data_ptr(), mask.data_ptr(), grad_out.data_ptr(),
                out.data_ptr(), grad_value.data_ptr(), grad_offset.data_ptr(), grad_mask.data_ptr(), n)
        _native.check(rc, "dcnv3_host_pipeline_run")

    def sync(self):
        with torch.cuda.device(self.device):
            _native.check(self._lib.dcnv3_host_pipeline_sync(self._h), "dcnv3_host_pipeline_sync")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            with torch.cuda.device(self.device):
                self._lib.dcnv3_host_pipeline_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
