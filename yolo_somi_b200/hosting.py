"""Hosting the DCNv3 layer in the reference's NCHW module zoo.

The reference bundles ``ops_dcnv3`` but never wires it into ``models/common.py`` / ``parse_model``
(SURVEY F1).  The zoo hosts its other deformable layer this way: a Conv-like wrapper (models/common.py:3768-3831),
a bottleneck that uses it as ``cv2`` (:3849-3859) and ``C3`` / ``C2f`` subclasses that swap their inner blocks
(:3862-3882).  This module does the same for the channels-last DCNv3 layer and REUSES the zoo's own building blocks:

    DCNv3_YOLO         NCHW in -> NHWC -> DCNv3 (this library) -> NCHW -> BatchNorm -> SiLU
    Bottleneck_DCNv3   cv1 = the zoo's 1x1 Conv, cv2 = DCNv3_YOLO, optional shortcut
    C3_DCNv3, C2f_DCNv3   subclasses of the zoo's C3 / C2f whose ``m`` holds Bottleneck_DCNv3 blocks

Inside the host tree ``Conv`` / ``C3`` / ``C2f`` are imported from ``models.common`` (so ``fuse()``, checkpoints and
``parse_model`` see the classes they know).  Outside it (this repo's own step harness, the GPU box) a minimal stand-in
with the same attribute names (``conv``/``bn``/``act``, ``cv1``/``cv2``/``cv3``/``m``) is used.  The patch that lists
these names in ``parse_model`` is ``integration/parse_model_dcnv3.patch``.
The permutes are views when the surrounding model runs in ``torch.channels_last`` memory format.
"""
from __future__ import annotations

import torch
from torch import nn

from .ops_dcnv3.modules import DCNv3


def _host_zoo():
    """(Conv, C3, C2f) of the host model zoo if this file is used inside it, else None."""
    try:
        from models import common as zoo           # the reference tree: models/common.py:55-66 (Conv), C3, C2f
        return zoo.Conv, zoo.C3, zoo.C2f
    except Exception:
        return None


def _standin_zoo():
    """Stand-ins for use outside the host tree: same constructor arguments and attribute names as the zoo's blocks."""

    class Conv(nn.Module):
        def __init__(self, c1, c2, k=1, s=1, p=None, g=1, d=1, act=True):
            super().__init__()
            pad = (d * (k - 1)) // 2 if p is None else p
            self.conv = nn.Conv2d(c1, c2, k, s, pad, dilation=d, groups=g, bias=False)
            self.bn = nn.BatchNorm2d(c2)
            self.act = nn.SiLU() if act is True else (act if isinstance(act, nn.Module) else nn.Identity())

        def forward(self, x):
            return self.act(self.bn(self.conv(x)))

    class C3(nn.Module):
        def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5):
            super().__init__()
            hidden = int(c2 * e)
            self.cv1, self.cv2, self.cv3 = Conv(c1, hidden), Conv(c1, hidden), Conv(2 * hidden, c2)
            self.m = nn.Sequential()

        def forward(self, x):
            return self.cv3(torch.cat((self.m(self.cv1(x)), self.cv2(x)), 1))

    class C2f(nn.Module):
        def __init__(self, c1, c2, n=1, shortcut=False, g=1, e=0.5):
            super().__init__()
            self.c = int(c2 * e)
            self.cv1, self.cv2 = Conv(c1, 2 * self.c), Conv((2 + n) * self.c, c2)
            self.m = nn.ModuleList()

        def forward(self, x):
            parts = list(self.cv1(x).split((self.c, self.c), 1))
            for block in self.m:
                parts.append(block(parts[-1]))
            return self.cv2(torch.cat(parts, 1))

    return Conv, C3, C2f


Conv, C3, C2f = _host_zoo() or _standin_zoo()


def groups_for(channels: int) -> int:
    """Group count for a DCNv3 layer of `channels`: group_channels 16 (else 32, else 8) -- the shapes the fastest
    kernels take -- preferring a multiple of 8 groups; otherwise the largest divisor that leaves <= 32 channels per
    group.  Always divides `channels` (the layer's constructor requires it, modules/dcnv3.py:252-254)."""
    for want_mult8 in (True, False):
        for gc in (16, 32, 8):
            if channels % gc == 0 and (not want_mult8 or (channels // gc) % 8 == 0):
                return channels // gc
    for gc in range(min(32, channels), 0, -1):
        if channels % gc == 0:
            return channels // gc
    return 1


class DCNv3_YOLO(nn.Module):
    """Conv-like wrapper of the DCNv3 layer for NCHW feature maps (c1 -> c2, stride 1).

    Stride 2 is rejected: the layer computes offsets / masks on the INPUT grid (its depthwise conv has stride 1,
    modules/dcnv3.py:276-289) while the core expects them on the OUTPUT grid, so the reference layer itself cannot run
    with stride != 1; downsample with the zoo's strided Conv in front instead."""

    def __init__(self, c1, c2, k=3, s=1, p=None, g=None, d=1, act=True):
        super().__init__()
        if s != 1:
            raise ValueError("DCNv3_YOLO: stride must be 1 (put a strided Conv in front); got %r" % (s,))
        self.pre = Conv(c1, c2, 1, 1) if c1 != c2 else nn.Identity()
        self.dcn = DCNv3(channels=c2, kernel_size=k, stride=1, pad=(d * (k - 1)) // 2 if p is None else p, dilation=d,
                         group=g or groups_for(c2))
        self.bn = nn.BatchNorm2d(c2)
        self.act = nn.SiLU() if act is True else (act if isinstance(act, nn.Module) else nn.Identity())

    def forward(self, x):
        x = self.pre(x).permute(0, 2, 3, 1)          # NCHW -> NHWC (a view under channels_last)
        x = self.dcn(x.contiguous()).permute(0, 3, 1, 2)
        return self.act(self.bn(x))


class Bottleneck_DCNv3(nn.Module):
    """The zoo's bottleneck (1x1 reduce, 3x3, optional residual) with the DCNv3 wrapper as the 3x3."""

    def __init__(self, c1, c2, shortcut=True, g=None, e=0.5):
        super().__init__()
        hidden = int(c2 * e)
        self.cv1 = Conv(c1, hidden, 1, 1)
        self.cv2 = DCNv3_YOLO(hidden, c2, 3, 1, g=g)
        self.add = shortcut and c1 == c2

    def forward(self, x):
        y = self.cv2(self.cv1(x))
        return x + y if self.add else y


class C3_DCNv3(C3):
    """The zoo's C3 with Bottleneck_DCNv3 inner blocks (the way C3_DCN swaps them, models/common.py:3862-3867)."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=None, e=0.5):
        super().__init__(c1, c2, n, shortcut, 1, e)
        hidden = int(c2 * e)
        self.m = nn.Sequential(*(Bottleneck_DCNv3(hidden, hidden, shortcut, g, e=1.0) for _ in range(n)))


class C2f_DCNv3(C2f):
    """The zoo's C2f with Bottleneck_DCNv3 inner blocks (models/common.py:3870-3882)."""

    def __init__(self, c1, c2, n=1, shortcut=False, g=None, e=0.5):
        super().__init__(c1, c2, n, shortcut, 1, e)
        self.m = nn.ModuleList(Bottleneck_DCNv3(self.c, self.c, shortcut, g, e=1.0) for _ in range(n))


# ----------------------------------------------------------------------------------------------------------------------
# Inference form (BASELINE configs[2]).  The reference folds every Conv's BatchNorm into the convolution before
# validation / detection (models/yolo.py `fuse()` over utils/torch_utils.py:202-222, called by attempt_load(fuse=True));
# the same folding applies to the BatchNorm behind a DCNv3 layer: it is an affine map per channel, so it goes into the
# layer's output_proj.  `GraphedInference` then replays the whole forward as one CUDA graph.
def _bn_scale_shift(bn: nn.BatchNorm2d):
    inv = torch.rsqrt(bn.running_var.double() + bn.eps)
    gamma = bn.weight.double() if bn.affine else torch.ones_like(inv)
    beta = bn.bias.double() if bn.affine else torch.zeros_like(inv)
    scale = gamma * inv
    return scale, beta - bn.running_mean.double() * scale


class BiasAct(nn.Module):
    """``act(x + bias)`` for NCHW feature maps, the tail of a folded Conv block (the reference's Conv.forward_fuse,
    models/common.py:55-66, with the folded bias taken out of the convolution).  16-bit channels-last CUDA tensors
    outside autograd go through one in-place pass of ``dcnv3_bias_act_sm100`` (csrc/dcnv3_hosting.cu); everything
    else through the plain PyTorch expression."""

    KINDS = {"identity": 0, "silu": 1}

    def __init__(self, bias: torch.Tensor, kind: str):
        super().__init__()
        if kind not in self.KINDS:
            raise ValueError("BiasAct: kind must be one of %s" % (tuple(self.KINDS),))
        self.kind = kind
        self.register_buffer("bias", bias.detach().clone().float())
        self._bias32 = None

    def _bias_f32(self):
        b = self.bias
        if b.dtype == torch.float32:
            return b
        if self._bias32 is None or self._bias32.device != b.device:       # model.half() narrowed the buffer
            self._bias32 = b.float()
        return self._bias32

    def forward(self, x):
        c = x.shape[1]
        if (x.is_cuda and x.dim() == 4 and x.dtype in (torch.float16, torch.bfloat16) and c % 8 == 0
                and not (torch.is_grad_enabled() and x.requires_grad) and x.permute(0, 2, 3, 1).is_contiguous()
                and x.data_ptr() % 16 == 0):
            from . import _native
            lib = _native.load()
            b = self._bias_f32()
            with torch.cuda.device(x.device):
                rc = lib.dcnv3_bias_act_sm100(x.data_ptr(), b.data_ptr(), x.data_ptr(), x.numel() // c, c, self.KINDS[self.kind],
                                              _native.F16 if x.dtype == torch.float16 else _native.BF16,
                                              torch.cuda.current_stream().cuda_stream)
            _native.check(rc, "dcnv3_bias_act_sm100")
            return x
        y = x + self.bias.to(x.dtype).view(1, -1, 1, 1)
        return nn.functional.silu(y) if self.kind == "silu" else y


@torch.no_grad()
def fuse_for_inference(model: nn.Module, half: bool = False) -> nn.Module:
    """Fold eval-mode BatchNorm layers into what feeds them, in place: ``conv`` + ``bn`` pairs of the zoo's Conv blocks
    (weight' = scale x weight, bias' = scale x bias + shift) and the BatchNorm of a ``DCNv3_YOLO`` wrapper into the
    DCNv3 layer's ``output_proj``.  The folded norms become ``nn.Identity``.  Where a Conv block's activation is SiLU or
    the identity, the folded bias leaves the convolution and joins the activation (``BiasAct``: one pass instead of
    ATen's broadcast add + SiLU).  ``half=True`` also narrows the model to fp16 as the reference's val / detect do
    (``model.half()``): no per-call weight casts under autocast.  Returns the model (in eval mode)."""
    model.eval()
    for m in model.modules():
        if isinstance(m, DCNv3_YOLO) and isinstance(m.bn, nn.BatchNorm2d):
            scale, shift = _bn_scale_shift(m.bn)
            lin = m.dcn.output_proj
            w = (lin.weight.double() * scale[:, None]).to(lin.weight.dtype)
            b = ((lin.bias.double() if lin.bias is not None else 0.0) * scale + shift).to(lin.weight.dtype)
            lin.weight.copy_(w)
            if lin.bias is None:
                lin.bias = nn.Parameter(b)
            else:
                lin.bias.copy_(b)
            m.bn = nn.Identity()
        elif (isinstance(getattr(m, "conv", None), nn.Conv2d) and isinstance(getattr(m, "bn", None), nn.BatchNorm2d)):
            scale, shift = _bn_scale_shift(m.bn)
            conv = m.conv
            w = (conv.weight.double() * scale[:, None, None, None]).to(conv.weight.dtype)
            b = ((conv.bias.double() if conv.bias is not None else 0.0) * scale + shift)
            conv.weight.copy_(w)
            act = getattr(m, "act", None)
            kind = "silu" if isinstance(act, nn.SiLU) else "identity" if isinstance(act, nn.Identity) else None
            if kind is not None:
                conv.bias = None
                m.act = BiasAct(b.to(conv.weight.device), kind)
            elif conv.bias is None:
                conv.bias = nn.Parameter(b.to(conv.weight.dtype).to(conv.weight.device))
            else:
                conv.bias.copy_(b.to(conv.weight.dtype))
            m.bn = nn.Identity()       # (the zoo's Conv.forward = act(bn(conv(x))) is its forward_fuse now)
    for p in model.parameters():
        p.requires_grad_(False)
    if half:
        model.half()
    return model


class GraphedInference:
    """The forward of an eval-mode model as ONE CUDA graph: ``y = GraphedInference(model, sample)(x)``.

    ``sample`` fixes shape, dtype and memory format; ``warmup`` eager forwards run first on the capture stream (cuDNN
    plans, tensor maps, the packed-weight caches of this library's fused producers -- frozen weights are packed once,
    so the graph holds no repack).  Each call copies ``x`` into the static input and replays; the returned tensors are
    the graph's static outputs (valid until the next call; clone to keep).  Weights must not change afterwards."""

    def __init__(self, model: nn.Module, sample: torch.Tensor, autocast_dtype: torch.dtype | None = torch.float16, warmup: int = 3):
        self.model = model.eval()
        self.dtype = autocast_dtype
        self._in = sample.clone()
        self._stream = torch.cuda.Stream(sample.device)
        self._stream.wait_stream(torch.cuda.current_stream(sample.device))
        with torch.cuda.stream(self._stream):
            for _ in range(warmup):
                self._forward()
        torch.cuda.current_stream(sample.device).wait_stream(self._stream)
        torch.cuda.synchronize(sample.device)
        self._g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._g, stream=self._stream):
            self._out = self._forward()

    def _forward(self):
        amp = torch.autocast("cuda", dtype=self.dtype) if self.dtype is not None else torch.autocast("cuda", enabled=False)
        with torch.no_grad(), amp:
            return self.model(self._in)

    def __call__(self, x: torch.Tensor):
        if x.shape != self._in.shape:
            raise ValueError("GraphedInference: static shape %s, got %s" % (tuple(self._in.shape), tuple(x.shape)))
        self._in.copy_(x, non_blocking=True)
        self._g.replay()
        return self._out
