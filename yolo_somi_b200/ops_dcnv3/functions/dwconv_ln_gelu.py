"""Fused producer of the DCNv3 layer's x1: depthwise conv + LayerNorm + GELU, channels-last, one
kernel (libdcnv3_sm100.so, ``dcnv3_dwconv_ln_gelu_sm100``) instead of the reference's
permute -> Conv2d(groups=C) -> permute -> LayerNorm -> GELU (models/ops_dcnv3/modules/dcnv3.py:
276-289,328-329).  The backward recomputes the unfused form under autograd (PyTorch kernels).
"""
from __future__ import annotations

import os
import weakref

import torch
import torch.nn.functional as F

from ... import _native

_DT = {torch.float16: _native.F16, torch.bfloat16: _native.BF16}


def eligible(x, conv, norm, act, dtype) -> bool:
    if os.environ.get("DCNV3_FUSED_DWCONV", "1") in ("0", ""):
        return False
    k = conv.kernel_size
    return (x.is_cuda and dtype in _DT and x.dim() == 4 and x.shape[-1] in (64, 128, 256)
            and isinstance(norm, torch.nn.LayerNorm) and isinstance(act, torch.nn.GELU)
            and getattr(act, "approximate", "none") == "none" and k[0] == k[1] and k[0] % 2 == 1 and k[0] <= 7
            and conv.stride == (1, 1) and conv.dilation == (1, 1) and conv.padding == ((k[0] - 1) // 2,) * 2
            and conv.groups == x.shape[-1] and conv.bias is not None and norm.elementwise_affine)


def _unfused(x, w, b, gamma, beta, eps):
    """The reference's sequence on a channels-last tensor (used by the backward)."""
    c = x.shape[-1]
    y = F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=(w.shape[-1] - 1) // 2, groups=c).permute(0, 2, 3, 1)
    return F.gelu(F.layer_norm(y, (c,), gamma, beta, eps))


_PACKED = {}   # id(conv weight) -> (weak refs, versions, dtype, packed tensors)


def _pack(w, b, gamma, beta, dtype):
    """Kernel-layout copies of the parameters ([k*k][C] taps in the I/O dtype, fp32 vectors),
    rebuilt only when one of the four tensors is another object or was modified in place."""
    tensors = (w, b, gamma, beta)
    # fingerprint: object identity, in-place version, storage pointer and device of all four tensors (`p.data = t`
    # and module.to(device) keep the Parameter object and its version).  `p.data.add_()` changes none of these, so
    # parameters that are being TRAINED are repacked on every call -- the cache only serves frozen weights.
    ver = tuple((t._version, t.data_ptr(), t.device) for t in tensors)
    training = torch.is_grad_enabled() and any(t.requires_grad for t in tensors)
    hit = _PACKED.get(id(w))
    if not training and hit is not None and hit[1] == ver and hit[2] == dtype and all(r() is t for r, t in zip(hit[0], tensors)):
        return hit[3]
    c, k = w.shape[0], w.shape[-1]
    packed = (w.detach().reshape(c, k * k).t().to(dtype).contiguous(),
              *(t.detach().float().contiguous() for t in (b, gamma, beta)))
    if len(_PACKED) > 64:
        _PACKED.clear()
    _PACKED[id(w)] = (tuple(weakref.ref(t) for t in tensors), ver, dtype, packed)
    return packed


class DwConvLnGelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, eps, dtype):
        lib = _native.load()
        n, h, wd, c = x.shape
        k = w.shape[-1]
        x16 = x.to(dtype).contiguous()
        wk, bf, gf, tf = _pack(w, b, gamma, beta, dtype)   # named: they outlive the launch
        out = torch.empty_like(x16)
        need_grad = any(ctx.needs_input_grad[:5])
        conv_out = torch.empty_like(x16) if need_grad else None     # pre-LayerNorm values for the backward
        with torch.cuda.device(x.device):
            rc = lib.dcnv3_dwconv_ln_gelu_sm100(
                x16.data_ptr(), wk.data_ptr(), bf.data_ptr(), gf.data_ptr(), tf.data_ptr(),
                out.data_ptr(), conv_out.data_ptr() if need_grad else None, n, h, wd, c, k, float(eps),
                _DT[dtype], torch.cuda.current_stream().cuda_stream)
        _native.check(rc, "dcnv3_dwconv_ln_gelu_sm100")
        ctx.save_for_backward(x16, w, b, gamma, beta, conv_out)
        ctx.eps, ctx.dtype, ctx.in_dtype = eps, dtype, x.dtype
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x16, w, b, gamma, beta, conv_out = ctx.saved_tensors
        c, k = x16.shape[-1], w.shape[-1]
        if k == 3 and os.environ.get("DCNV3_FUSED_DWCONV_BWD", "1") not in ("0", ""):
            # two passes (csrc/dcnv3_dwconv_bwd.cu): LayerNorm + GELU backward with the channel sums, then the depthwise
            # convolution's dgrad + wgrad on TMA-staged windows
            lib = _native.load()
            n, h, wd, _ = x16.shape
            wk, _, gf, tf = _pack(w, b, gamma, beta, ctx.dtype)
            g16 = g.to(ctx.dtype).contiguous()
            du = torch.empty_like(x16)
            g_x = torch.empty_like(x16)
            params = torch.empty(12 * c, dtype=torch.float32, device=x16.device)
            with torch.cuda.device(x16.device):
                rc = lib.dcnv3_dwconv_ln_gelu_backward_sm100(
                    x16.data_ptr(), conv_out.data_ptr(), g16.data_ptr(), wk.data_ptr(), gf.data_ptr(), tf.data_ptr(),
                    du.data_ptr(), g_x.data_ptr(), params.data_ptr(), n, h, wd, c, k, float(ctx.eps), _DT[ctx.dtype],
                    torch.cuda.current_stream().cuda_stream)
            _native.check(rc, "dcnv3_dwconv_ln_gelu_backward_sm100")
            g_w = params[:9 * c].reshape(9, c).t().reshape(c, 1, 3, 3)
            g_b, g_gamma, g_beta = params[9 * c:10 * c], params[10 * c:11 * c], params[11 * c:]
            return (g_x.to(ctx.in_dtype), g_w.to(w.dtype), g_b.to(b.dtype), g_gamma.to(gamma.dtype),
                    g_beta.to(beta.dtype), None, None)
        # (other kernel sizes) LayerNorm + GELU re-derived from the saved convolution output under autograd ...
        with torch.enable_grad():
            y = conv_out.detach().requires_grad_(True)
            gl, bl = gamma.detach().to(ctx.dtype).requires_grad_(True), beta.detach().to(ctx.dtype).requires_grad_(True)
            z = F.gelu(F.layer_norm(y, (c,), gl, bl, ctx.eps))
            g_y, g_gamma, g_beta = torch.autograd.grad(z, (y, gl, bl), g.to(z.dtype))
        # ... and the convolution's own backward, without re-running its forward
        g_x, g_w, g_b = torch.ops.aten.convolution_backward(
            g_y.permute(0, 3, 1, 2), x16.permute(0, 3, 1, 2), w.detach().to(ctx.dtype), [c], [1, 1],
            [(k - 1) // 2] * 2, [1, 1], False, [0, 0], c, [True, True, True])
        return (g_x.permute(0, 2, 3, 1).to(ctx.in_dtype), g_w.to(w.dtype), g_b.to(b.dtype), g_gamma.to(gamma.dtype),
                g_beta.to(beta.dtype), None, None)


def dwconv_ln_gelu(x, conv, norm, dtype):
    """x1 [N,H,W,C] in ``dtype`` from the channels-last input x."""
    return DwConvLnGelu.apply(x, conv.weight, conv.bias, norm.weight, norm.bias, norm.eps, dtype)
