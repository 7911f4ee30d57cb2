"""YOLOv5l-DCNv3: the model BASELINE.json configs[2] / [3] are quoted on, for the step harness (train_step_ddp.py).

The reference's flagship yaml cannot be built as shipped (it names a class that does not exist) and nothing
in its zoo names DCNv3 (F1), so the graph is stated here explicitly: the stock YOLOv5l layout (depth 1.0, width 1.0:
models/yolov5l.yaml upstream -- Conv / C3 / SPPF backbone, PANet head, Detect on P3 / P4 / P5) with the four C3 stages
of the HEAD replaced by ``C3_DCNv3`` (hosting.py), i.e. twelve DCNv3 layers:

    stage  stride  map @640   DCNv3 channels  groups (16 ch each)
    13     16      40 x 40    256             16
    17      8      80 x 80    128              8
    20     16      40 x 40    256             16
    23     32      20 x 20    512             32

`layers()` returns that list (index, from, module, arguments) and bench.py prints it, so "YOLOv5l-DCNv3" is
reproducible from the JSON line alone.  Plain PyTorch (cuDNN / cuBLAS) everywhere except the DCNv3 layers, whose core
and fused producers are this library's sm_100a kernels.  nc = 10 (data/VisDrone.yaml:17), 3 anchors per level.
"""
from __future__ import annotations

import torch
from torch import nn

from .hosting import C3, C3_DCNv3, Conv

# (from, module, args) in the zoo's yaml convention; c1 is implied by `from`
_SPEC = [
    (-1, "Conv", (64, 6, 2, 2)),          # 0  P1/2
    (-1, "Conv", (128, 3, 2)),            # 1  P2/4
    (-1, "C3", (128, 3)),                 # 2
    (-1, "Conv", (256, 3, 2)),            # 3  P3/8
    (-1, "C3", (256, 6)),                 # 4
    (-1, "Conv", (512, 3, 2)),            # 5  P4/16
    (-1, "C3", (512, 9)),                 # 6
    (-1, "Conv", (1024, 3, 2)),           # 7  P5/32
    (-1, "C3", (1024, 3)),                # 8
    (-1, "SPPF", (1024, 5)),              # 9
    (-1, "Conv", (512, 1, 1)),            # 10
    (-1, "Upsample", (2,)),               # 11
    ((-1, 6), "Concat", ()),              # 12
    (-1, "C3_DCNv3", (512, 3, False)),    # 13
    (-1, "Conv", (256, 1, 1)),            # 14
    (-1, "Upsample", (2,)),               # 15
    ((-1, 4), "Concat", ()),              # 16
    (-1, "C3_DCNv3", (256, 3, False)),    # 17  P3/8
    (-1, "Conv", (256, 3, 2)),            # 18
    ((-1, 14), "Concat", ()),             # 19
    (-1, "C3_DCNv3", (512, 3, False)),    # 20  P4/16
    (-1, "Conv", (512, 3, 2)),            # 21
    ((-1, 10), "Concat", ()),             # 22
    (-1, "C3_DCNv3", (1024, 3, False)),   # 23  P5/32
    ((17, 20, 23), "Detect", ()),         # 24
]


def layers():
    """The layer list as plain data (for the bench's JSON line)."""
    return [[i, list(f) if isinstance(f, tuple) else f, m, list(a)] for i, (f, m, a) in enumerate(_SPEC)]


class SPPF(nn.Module):
    """Spatial pyramid pooling, fast form: three chained 5x5 max-pools concatenated with the input."""

    def __init__(self, c1, c2, k=5):
        super().__init__()
        self.cv1, self.cv2 = Conv(c1, c1 // 2, 1, 1), Conv(c1 * 2, c2, 1, 1)
        self.m = nn.MaxPool2d(k, 1, k // 2)

    def forward(self, x):
        x = self.cv1(x)
        y1 = self.m(x)
        y2 = self.m(y1)
        return self.cv2(torch.cat((x, y1, y2, self.m(y2)), 1))


class _PlainC3(C3):
    """The zoo's C3 with its standard bottlenecks (1x1 Conv, 3x3 Conv, residual)."""

    class _Bottleneck(nn.Module):
        def __init__(self, c, shortcut):
            super().__init__()
            self.cv1, self.cv2, self.add = Conv(c, c, 1, 1), Conv(c, c, 3, 1), shortcut

        def forward(self, x):
            y = self.cv2(self.cv1(x))
            return x + y if self.add else y

    def __init__(self, c1, c2, n=1, shortcut=True):
        super().__init__(c1, c2, n, shortcut)
        if len(self.m) == 0:       # the stand-in zoo leaves `m` empty; the host zoo has already filled it
            self.m = nn.Sequential(*(self._Bottleneck(c2 // 2, shortcut) for _ in range(n)))


class Detect(nn.Module):
    """Detection head: one 1x1 convolution per level to na * (nc + 5) channels; returns the raw maps
    [B, na, H, W, nc + 5] (training form of the zoo's Detect)."""

    def __init__(self, nc, ch, na=3):
        super().__init__()
        self.nc, self.na, self.no = nc, na, nc + 5
        self.m = nn.ModuleList(nn.Conv2d(c, self.no * na, 1) for c in ch)

    def forward(self, xs):
        out = []
        for x, conv in zip(xs, self.m):
            y = conv(x)
            b, _, h, w = y.shape
            out.append(y.view(b, self.na, self.no, h, w).permute(0, 1, 3, 4, 2))
        return out


class YOLOv5lDCNv3(nn.Module):
    def __init__(self, nc=10, ch=3):
        super().__init__()
        mods, outs = [], []
        self.routes = []
        for i, (f, name, a) in enumerate(_SPEC):
            srcs = [f] if isinstance(f, int) else list(f)
            cin = [ch if (s == -1 and i == 0) else outs[s if s >= 0 else i + s] for s in srcs]
            if name == "Conv":
                m, cout = Conv(cin[0], *a), a[0]
            elif name == "C3":
                m, cout = _PlainC3(cin[0], a[0], a[1]), a[0]
            elif name == "C3_DCNv3":
                m, cout = C3_DCNv3(cin[0], a[0], a[1], a[2]), a[0]
            elif name == "SPPF":
                m, cout = SPPF(cin[0], *a), a[0]
            elif name == "Upsample":
                m, cout = nn.Upsample(scale_factor=a[0], mode="nearest"), cin[0]
            elif name == "Concat":
                m, cout = None, sum(cin)
            else:
                m, cout = Detect(nc, cin), 0
            mods.append(m if m is not None else nn.Identity())
            outs.append(cout)
            self.routes.append(srcs)
        self.model = nn.ModuleList(mods)
        self.nc = nc
        self.keep = {s if s >= 0 else i + s for i, srcs in enumerate(self.routes) for s in srcs if s != -1}

    def forward(self, x):
        saved = {}
        for i, (m, srcs, (_, name, _a)) in enumerate(zip(self.model, self.routes, _SPEC)):
            if name == "Concat":
                x = torch.cat([x if s == -1 else saved[s if s >= 0 else i + s] for s in srcs], 1)
            elif name == "Detect":
                x = m([saved[s] for s in srcs])
            else:
                x = m(x)
            if i in self.keep:
                saved[i] = x
        return x

    def dcnv3_layers(self):
        from .ops_dcnv3.modules import DCNv3
        return [m for m in self.modules() if isinstance(m, DCNv3)]
