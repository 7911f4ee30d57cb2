"""Hosting the DCNv3 layer in the reference's NCHW module zoo.

The reference bundles ``ops_dcnv3`` but never wires it into ``models/common.py`` / ``parse_model``
(SURVEY F1).  These modules follow the way the zoo hosts DCNv2 -- a Conv-like wrapper
(models/common.py:3768-3831), a bottleneck using it as ``cv2`` (:3849-3859) and the ``C3`` / ``C2f``
containers built from it (:3862-3882) -- for the channels-last DCNv3 layer:

    DCNv3_YOLO         NCHW in -> NHWC -> DCNv3 (this library) -> NCHW -> BatchNorm -> SiLU
    Bottleneck_DCNv3   cv1 = 1x1 Conv, cv2 = DCNv3_YOLO, optional shortcut
    C3_DCNv3, C2f_DCNv3   the zoo's CSP containers with that bottleneck

To use them from a model yaml, import the names into ``models/yolo.py`` and add them to the
channel-handling lists of ``parse_model`` (models/yolo.py:1472-1492), see INTEGRATION.md.  The
permutes are views when the surrounding model runs in ``torch.channels_last`` memory format.
"""
from __future__ import annotations

import torch
from torch import nn

from .ops_dcnv3.modules import DCNv3


def autopad(k, p=None, d=1):
    if d > 1:
        k = d * (k - 1) + 1 if isinstance(k, int) else [d * (x - 1) + 1 for x in k]
    if p is None:
        p = k // 2 if isinstance(k, int) else [x // 2 for x in k]
    return p


class Conv(nn.Module):
    """The zoo's standard convolution block: Conv2d + BatchNorm2d + SiLU."""
    default_act = nn.SiLU()

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, d=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, autopad(k, p, d), groups=g, dilation=d, bias=False)
        self.bn = nn.BatchNorm2d(c2)
        self.act = self.default_act if act is True else act if isinstance(act, nn.Module) else nn.Identity()

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


def _groups_for(channels: int) -> int:
    """16 channels per group (the fast kernels' shape), at least one group."""
    return max(1, channels // 16)


class DCNv3_YOLO(nn.Module):
    """Conv-like wrapper of the DCNv3 layer for NCHW feature maps (c1 -> c2, stride 1 or 2)."""

    def __init__(self, c1, c2, k=3, s=1, p=None, g=None, d=1, act=True):
        super().__init__()
        self.pre = Conv(c1, c2, 1, 1) if c1 != c2 else nn.Identity()
        self.dcn = DCNv3(channels=c2, kernel_size=k, stride=s, pad=autopad(k, p, d), dilation=d,
                         group=g or _groups_for(c2))
        self.bn = nn.BatchNorm2d(c2)
        self.act = Conv.default_act if act is True else act if isinstance(act, nn.Module) else nn.Identity()

    def forward(self, x):
        x = self.pre(x).permute(0, 2, 3, 1)          # NCHW -> NHWC (a view under channels_last)
        x = self.dcn(x.contiguous()).permute(0, 3, 1, 2)
        return self.act(self.bn(x))


class Bottleneck_DCNv3(nn.Module):
    def __init__(self, c1, c2, shortcut=True, g=None, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = DCNv3_YOLO(c_, c2, 3, 1, g=g)
        self.add = shortcut and c1 == c2

    def forward(self, x):
        return x + self.cv2(self.cv1(x)) if self.add else self.cv2(self.cv1(x))


class C3_DCNv3(nn.Module):
    """CSP bottleneck with 3 convolutions (the zoo's C3) whose inner blocks are Bottleneck_DCNv3."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=None, e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(Bottleneck_DCNv3(c_, c_, shortcut, g, e=1.0) for _ in range(n)))

    def forward(self, x):
        return self.cv3(torch.cat((self.m(self.cv1(x)), self.cv2(x)), 1))


class C2f_DCNv3(nn.Module):
    def __init__(self, c1, c2, n=1, shortcut=False, g=None, e=0.5):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneck_DCNv3(self.c, self.c, shortcut, g, e=1.0) for _ in range(n))

    def forward(self, x):
        y = list(self.cv1(x).split((self.c, self.c), 1))
        y.extend(m(y[-1]) for m in self.m)
        return self.cv2(torch.cat(y, 1))
