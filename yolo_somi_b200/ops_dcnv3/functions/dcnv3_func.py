"""``DCNv3Function``: autograd wrapper of the sm_100a DCNv3 core.

Mirrors the reference operator (models/ops_dcnv3/functions/dcnv3_func.py:19-89): same name, same
15 positional arguments to ``apply`` (input, offset, mask, kernel_h, kernel_w, stride_h, stride_w,
pad_h, pad_w, dilation_h, dilation_w, group, group_channels, offset_scale, im2col_step), same
three gradients + twelve ``None`` from ``backward``, not twice differentiable, autocast-transparent
(runs in whatever dtype autocast handed it), same ONNX node (``mmdeploy::TRTDCNv3``).

The reference's ``dcnv3_core_pytorch`` (a slow CPU/grid_sample implementation "for debug and test
only", :147-188) is intentionally NOT part of the product: it lives on as the oracle in
``oracle/dcnv3_oracle.py`` and only tests / the CPU baseline use it.
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import dcnv3_ext

_GEOM_FIELDS = ("kernel_h", "kernel_w", "stride_h", "stride_w", "pad_h", "pad_w",
                "dilation_h", "dilation_w", "group", "group_channels", "offset_scale")


class DCNv3Function(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                dilation_h, dilation_w, group, group_channels, offset_scale, im2col_step):
        geom = (kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
                group, group_channels, offset_scale)
        for name, val in zip(_GEOM_FIELDS, geom):  # same ctx attributes as the reference (:27-38)
            setattr(ctx, name, val)
        ctx.im2col_step = im2col_step
        ctx.save_for_backward(input, offset, mask)
        return dcnv3_ext.dcnv3_forward(input, offset, mask, *geom, im2col_step)

    @staticmethod
    @once_differentiable
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        input, offset, mask = ctx.saved_tensors
        geom = tuple(getattr(ctx, name) for name in _GEOM_FIELDS)
        grads = dcnv3_ext.dcnv3_backward(input, offset, mask, *geom, grad_output.contiguous(),
                                         ctx.im2col_step)
        return (*grads, *(None,) * 12)

    @staticmethod
    def symbolic(g, input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                 dilation_h, dilation_w, group, group_channels, offset_scale, im2col_step):
        """ONNX export node, identical to the reference's (dcnv3_func.py:63-89)."""
        ints = dict(kernel_h=kernel_h, kernel_w=kernel_w, stride_h=stride_h, stride_w=stride_w,
                    pad_h=pad_h, pad_w=pad_w, dilation_h=dilation_h, dilation_w=dilation_w,
                    group=group, group_channels=group_channels, im2col_step=im2col_step)
        attrs = {f"{k}_i": int(v) for k, v in ints.items()}
        attrs["offset_scale_f"] = float(offset_scale)
        return g.op("mmdeploy::TRTDCNv3", input, offset, mask, **attrs)


def dcnv3_core(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
               dilation_h, dilation_w, group, group_channels, offset_scale, im2col_step=256):
    """Functional form of ``DCNv3Function.apply`` with the argument order of the reference's
    ``dcnv3_core_pytorch`` (im2col_step optional)."""
    return DCNv3Function.apply(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w,
                               pad_h, pad_w, dilation_h, dilation_w, group, group_channels,
                               offset_scale, im2col_step)


def dcnv3_core_pytorch(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                       dilation_h, dilation_w, group, group_channels, offset_scale):
    """The NAME of the reference's debug implementation (functions/dcnv3_func.py:147-188), with its signature, so that
    ``from models.ops_dcnv3.functions import DCNv3Function, dcnv3_core_pytorch`` keeps importing.  Here it computes the
    same function on the sm_100a kernels (CUDA tensors only: this library has no PyTorch / CPU path; the grid_sample
    restatement lives in ``oracle/`` as test infrastructure and is never imported from here)."""
    return dcnv3_core(input, offset, mask, kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w,
                      dilation_h, dilation_w, group, group_channels, offset_scale)
