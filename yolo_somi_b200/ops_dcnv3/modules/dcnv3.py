"""``DCNv3`` layer (channels-last in, channels-last out) on the sm_100a core.

Mirrors the reference module (models/ops_dcnv3/modules/dcnv3.py:222-379): identical constructor
signature, identical parameter names/shapes (``dw_conv.0.*``, ``dw_conv.1.1.*``, ``offset.*``,
``mask.*``, ``input_proj.*``, ``output_proj.*``, optional ``center_feature_scale_proj_*``) so
checkpoints move across unchanged, identical initialisation (:307-315) and forward semantics
(:317-379):

    x      = input_proj(input)
    x1     = GELU(LN(dwconv3x3(input)))                      (norm/act selectable)
    offset = offset(x1);  mask = softmax_P(mask(x1))
    y      = DCNv3Function(x, offset, mask, k, k, s, s, p, p, d, d, G, C/G, offset_scale, 256)
    y      = y*(1-cfs) + x*cfs                               (only if center_feature_scale)
    out    = output_proj(y)
"""
from __future__ import annotations

import warnings

import torch
import torch.nn.functional as F
from torch import nn

from ..functions import DCNv3Function
from ..functions import dwconv_ln_gelu, offset_mask_proj


class to_channels_first(nn.Module):
    def forward(self, x):
        return x.permute(0, 3, 1, 2)


class to_channels_last(nn.Module):
    def forward(self, x):
        return x.permute(0, 2, 3, 1)


def build_norm_layer(dim, norm_layer, in_format="channels_last", out_format="channels_last",
                     eps=1e-6):
    """'BN' | 'LN' with layout adapters; same module indices as the reference (:41-62) so that
    ``dw_conv.1.1`` is the norm when the input is channels-first."""
    if norm_layer not in ("BN", "LN"):
        raise NotImplementedError(f"build_norm_layer does not support {norm_layer}")
    native = "channels_first" if norm_layer == "BN" else "channels_last"
    norm = nn.BatchNorm2d(dim) if norm_layer == "BN" else nn.LayerNorm(dim, eps=eps)
    flip = {"channels_first": to_channels_first, "channels_last": to_channels_last}
    seq = []
    if in_format != native:
        seq.append(flip[native]())
    seq.append(norm)
    if out_format != native:
        seq.append(flip[out_format]())
    return nn.Sequential(*seq)


def build_act_layer(act_layer):
    table = {"ReLU": lambda: nn.ReLU(inplace=True), "SiLU": lambda: nn.SiLU(inplace=True),
             "GELU": nn.GELU}
    if act_layer not in table:
        raise NotImplementedError(f"build_act_layer does not support {act_layer}")
    return table[act_layer]()


def _is_power_of_2(n):
    if not isinstance(n, int) or n < 0:
        raise ValueError(f"invalid input for _is_power_of_2: {n} (type: {type(n)})")
    return n != 0 and (n & (n - 1)) == 0


class _TallLinear(torch.autograd.Function):
    """``F.linear`` for tall 16-bit activations [M, C] with M >> C: identical forward; the backward
    takes the bias gradient as a ones-row GEMM instead of autograd's dim-0 reduction kernel, which at
    the layer's M = N*H*W = 102,400 costs 100-300 us per projection (profiles/README.md, r1_v4).

    Under autocast (fp32 parameters, 16-bit activations: the training step's mode) the operands are cast
    to the autocast dtype here, as autocast's own ``linear`` does, and the gradients go back in the
    parameters' dtype -- ``nn.Linear`` under autocast would take autograd's reduction for the bias."""

    @staticmethod
    def forward(ctx, x, weight, bias, dtype):
        with torch.autocast(x.device.type, enabled=False):
            x16, w16 = x.to(dtype), weight.to(dtype)
            ctx.save_for_backward(x16, w16)
            ctx.dtypes = (x.dtype, weight.dtype, bias.dtype)
            return F.linear(x16, w16, bias.to(dtype))

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        xd, wd, bd = ctx.dtypes
        g2, x2 = g.reshape(-1, g.shape[-1]).to(x.dtype), x.reshape(-1, x.shape[-1])
        gx = (g2 @ weight).reshape(x.shape).to(xd) if ctx.needs_input_grad[0] else None
        gw = (g2.t() @ x2).to(wd) if ctx.needs_input_grad[1] else None
        gb = None
        if ctx.needs_input_grad[2]:
            gb = offset_mask_proj.column_sums(g2).to(bd)
        return gx, gw, gb, None


def _linear(x, lin):
    if x.is_cuda and lin.bias is not None and torch.is_grad_enabled() and x.numel() // x.shape[-1] >= 4096:
        amp = torch.is_autocast_enabled("cuda")
        dtype = torch.get_autocast_dtype("cuda") if amp else x.dtype
        if dtype in (torch.float16, torch.bfloat16) and (amp or lin.weight.dtype == x.dtype):
            return _TallLinear.apply(x, lin.weight, lin.bias, dtype)
    return lin(x)


class CenterFeatureScaleModule(nn.Module):
    def forward(self, query, center_feature_scale_proj_weight, center_feature_scale_proj_bias):
        return F.linear(query, center_feature_scale_proj_weight,
                        center_feature_scale_proj_bias).sigmoid()


class DCNv3(nn.Module):
    def __init__(self, channels=64, kernel_size=3, dw_kernel_size=None, stride=1, pad=1,
                 dilation=1, group=4, offset_scale=1.0, act_layer="GELU", norm_layer="LN",
                 center_feature_scale=False, use_dcn_v4_op=False):
        super().__init__()
        if channels % group != 0:
            raise ValueError(f"channels must be divisible by group, but got {channels} and {group}")
        if use_dcn_v4_op:
            # the reference defers to an external DCNv4 package here (modules/dcnv3.py:17-20,344-368)
            raise NotImplementedError("use_dcn_v4_op=True needs the external DCNv4 extension")
        if not _is_power_of_2(channels // group):
            warnings.warn("channels // group is not a power of 2: the DCNv3 core falls back to its "
                          "generic (slower) kernels")
        dw_kernel_size = kernel_size if dw_kernel_size is None else dw_kernel_size
        self.channels, self.group, self.group_channels = channels, group, channels // group
        self.kernel_size, self.dw_kernel_size = kernel_size, dw_kernel_size
        self.stride, self.pad, self.dilation = stride, pad, dilation
        self.offset_scale = offset_scale
        self.center_feature_scale = center_feature_scale
        self.use_dcn_v4_op = use_dcn_v4_op

        pts = group * kernel_size * kernel_size
        self.dw_conv = nn.Sequential(
            nn.Conv2d(channels, channels, dw_kernel_size, stride=1,
                      padding=(dw_kernel_size - 1) // 2, groups=channels),
            build_norm_layer(channels, norm_layer, "channels_first", "channels_last"),
            build_act_layer(act_layer))
        self.offset = nn.Linear(channels, pts * 2)
        self.mask = nn.Linear(channels, pts)
        self.input_proj = nn.Linear(channels, channels)
        self.output_proj = nn.Linear(channels, channels)
        self._reset_parameters()
        if center_feature_scale:
            self.center_feature_scale_proj_weight = nn.Parameter(
                torch.zeros((group, channels), dtype=torch.float))
            self.center_feature_scale_proj_bias = nn.Parameter(
                torch.zeros((group,), dtype=torch.float))
            self.center_feature_scale_module = CenterFeatureScaleModule()

    def _reset_parameters(self):
        # offset/mask start at zero: the layer is a 3x3 average pool at init (modules/dcnv3.py:307-315)
        for lin in (self.offset, self.mask):
            nn.init.zeros_(lin.weight)
            nn.init.zeros_(lin.bias)
        for lin in (self.input_proj, self.output_proj):
            nn.init.xavier_uniform_(lin.weight)
            nn.init.zeros_(lin.bias)

    def forward(self, input):
        """input, output: (N, H, W, C)."""
        n, h, w, _ = input.shape
        x = _linear(input, self.input_proj)
        conv, norm, act = self.dw_conv[0], self.dw_conv[1][-1], self.dw_conv[2]
        if dwconv_ln_gelu.eligible(input, conv, norm, act, x.dtype):
            # depthwise conv + LayerNorm + GELU in one channels-last pass (csrc/dcnv3_dwconv.cu)
            x1 = dwconv_ln_gelu.dwconv_ln_gelu(input, conv, norm, x.dtype)
        else:
            x1 = self.dw_conv(input.permute(0, 3, 1, 2))
        if offset_mask_proj.eligible(x1, self.group, self.kernel_size * self.kernel_size, x.dtype):
            # one tcgen05 GEMM with bias, softmax and cast in its epilogue (csrc/dcnv3_proj.cu)
            offset, mask = offset_mask_proj.offset_mask_proj(x1, self.offset, self.mask, self.group, x.dtype)
        else:
            offset = self.offset(x1)
            mask = F.softmax(self.mask(x1).reshape(n, h, w, self.group, -1), -1)
            mask = mask.reshape(n, h, w, -1).type(x.dtype)
        k, s, p, d = self.kernel_size, self.stride, self.pad, self.dilation
        y = DCNv3Function.apply(x, offset, mask, k, k, s, s, p, p, d, d, self.group,
                                self.group_channels, self.offset_scale, 256)
        if self.center_feature_scale:
            cfs = self.center_feature_scale_module(x1, self.center_feature_scale_proj_weight,
                                                   self.center_feature_scale_proj_bias)
            cfs = cfs.repeat_interleave(self.group_channels, dim=-1)  # per group -> per channel
            y = y * (1 - cfs) + x * cfs
        return _linear(y, self.output_proj)


class DCNv3_pytorch(DCNv3):
    """The NAME of the reference's debug layer (modules/dcnv3.py:95-219): same constructor, same parameters, same
    ``state_dict`` keys as ``DCNv3``, so imports (``from models.ops_dcnv3.modules import DCNv3, DCNv3_pytorch``) and
    checkpoints that pickled this class path keep loading.  Here it IS the sm_100a layer -- there is no PyTorch / CPU
    sampler in this library (the reference's pure-PyTorch form is restated in ``oracle/`` as test infrastructure)."""
