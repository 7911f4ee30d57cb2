"""Seeded input generators shared by make_golden.py (which runs the REAL reference in the build
container) and by the tests (which replay the same inputs through the oracle / CUDA path).

Inputs are regenerated from numpy's PCG64 stream instead of being stored; every fixture carries a
float64 checksum of its inputs so a drifting generator is detected, not silently accepted.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict

import numpy as np


@dataclass(frozen=True)
class Case:
    name: str
    N: int
    H: int
    W: int
    G: int
    gc: int
    kh: int = 3
    kw: int = 3
    sh: int = 1
    sw: int = 1
    ph: int = 1
    pw: int = 1
    dh: int = 1
    dw: int = 1
    sigma: float = 1.0
    dist: str = "normal"        # "normal" | "reftest"
    grad: str = "normal"        # "normal" | "ones"
    seed: int = 0

    @property
    def Ho(self):
        return (self.H + 2 * self.ph - (self.dh * (self.kh - 1) + 1)) // self.sh + 1

    @property
    def Wo(self):
        return (self.W + 2 * self.pw - (self.dw * (self.kw - 1) + 1)) // self.sw + 1

    @property
    def P(self):
        return self.kh * self.kw

    @property
    def geom(self):
        """Positional tail of dcnv3_core_pytorch / DCNv3Function (without im2col_step)."""
        return (self.kh, self.kw, self.sh, self.sw, self.ph, self.pw, self.dh, self.dw,
                self.G, self.gc, self.sigma)

    def asdict(self):
        return asdict(self)


def make_inputs(c: Case):
    """float64 numpy (value, offset, mask, grad_out) for a case."""
    rng = np.random.default_rng(1000 + c.seed)
    C = c.G * c.gc
    if c.dist == "reftest":
        # distributions of the reference's own script, models/ops_dcnv3/test.py:35-39
        value = rng.random((c.N, c.H, c.W, C)) * 0.01
        offset = rng.random((c.N, c.Ho, c.Wo, c.G * c.P * 2)) * 10
        mask = rng.random((c.N, c.Ho, c.Wo, c.G, c.P)) + 1e-5
        mask = mask / mask.sum(-1, keepdims=True)
    else:
        value = rng.standard_normal((c.N, c.H, c.W, C))
        offset = rng.standard_normal((c.N, c.Ho, c.Wo, c.G * c.P * 2)) * 1.5
        logits = rng.standard_normal((c.N, c.Ho, c.Wo, c.G, c.P))
        e = np.exp(logits - logits.max(-1, keepdims=True))
        mask = e / e.sum(-1, keepdims=True)
    mask = mask.reshape(c.N, c.Ho, c.Wo, c.G * c.P)
    if c.grad == "ones":
        grad_out = np.ones((c.N, c.Ho, c.Wo, C))
    else:
        grad_out = rng.standard_normal((c.N, c.Ho, c.Wo, C))
    return value, offset, mask, grad_out


def input_checksum(arrs) -> float:
    return float(sum(np.asarray(a, dtype=np.float64).sum() for a in arrs))


# ---- the reference script's own cases: models/ops_dcnv3/test.py:19-30,93-216,257-260
REFTEST_FWD = Case("reftest_fwd", N=2, H=8, W=8, G=4, gc=16, sigma=2.0, dist="reftest",
                   grad="ones", seed=3)
REFTEST_BWD = [Case(f"reftest_bwd_gc{d}", N=2, H=8, W=8, G=2, gc=d, sigma=2.0, dist="reftest",
                    grad="ones", seed=30 + i)
               for i, d in enumerate((1, 16, 30, 32, 64, 71, 1025))]

# ---- (K, stride, pad, dilation, sigma) sweep on non-square maps; nothing in the reference pins
#      these, the golden values come from running dcnv3_core_pytorch itself
_SWEEP = [  # kh kw  s  pad d  sigma
    (3, 3, 1, 1, 1, 1.0),
    (3, 3, 1, 1, 1, 2.5),
    (3, 3, 2, 1, 1, 1.0),
    (3, 3, 1, 2, 2, 1.5),
    (5, 5, 1, 2, 1, 1.0),
    (3, 3, 2, 0, 1, 1.0),
    (3, 5, 1, 1, 1, 1.0),   # non-square kernel, equal pads
    (1, 1, 1, 0, 1, 1.0),
]
SWEEP = [Case(f"sweep_k{kh}x{kw}_s{s}_p{p}_d{d}_sig{sig}", N=2, H=9, W=13, G=2, gc=4,
              kh=kh, kw=kw, sh=s, sw=s, ph=p, pw=p, dh=d, dw=d, sigma=sig, seed=100 + i)
         for i, (kh, kw, s, p, d, sig) in enumerate(_SWEEP)]

# ---- BASELINE.json configs[0]: the reference's CPU-runnable correctness case
CFG1 = Case("cfg1_n2_40x40_c64_g4", N=2, H=40, W=40, G=4, gc=16, seed=7)

ALL = [REFTEST_FWD] + REFTEST_BWD + SWEEP + [CFG1]
BY_NAME = {c.name: c for c in ALL}

# elements of big tensors stored in the fixture: every STRIDE-th (plus the float64 sum)
SAMPLE_STRIDE = 13
BIG = 60_000


# ---- module-level fixture: DCNv3_pytorch (models/ops_dcnv3/modules/dcnv3.py:95-219)
@dataclass(frozen=True)
class ModuleCase:
    name: str
    channels: int = 64
    group: int = 4
    kernel_size: int = 3
    stride: int = 1
    pad: int = 1
    dilation: int = 1
    offset_scale: float = 1.0
    center_feature_scale: bool = False
    N: int = 2
    H: int = 12
    W: int = 10
    seed: int = 0


def make_module_state(m: ModuleCase):
    """Seeded float32 parameters under the reference's state_dict keys, and a float32 input."""
    rng = np.random.default_rng(5000 + m.seed)
    C, G, P = m.channels, m.group, m.kernel_size ** 2
    f = lambda *s, scale=1.0: (rng.standard_normal(s) * scale).astype(np.float32)
    state = {
        "dw_conv.0.weight": f(C, 1, m.kernel_size, m.kernel_size, scale=0.3),
        "dw_conv.0.bias": f(C, scale=0.1),
        "dw_conv.1.1.weight": (1.0 + 0.1 * rng.standard_normal(C)).astype(np.float32),
        "dw_conv.1.1.bias": f(C, scale=0.1),
        "offset.weight": f(G * P * 2, C, scale=0.15),
        "offset.bias": f(G * P * 2, scale=0.5),
        "mask.weight": f(G * P, C, scale=0.2),
        "mask.bias": f(G * P, scale=0.2),
        "input_proj.weight": f(C, C, scale=C ** -0.5),
        "input_proj.bias": f(C, scale=0.1),
        "output_proj.weight": f(C, C, scale=C ** -0.5),
        "output_proj.bias": f(C, scale=0.1),
    }
    if m.center_feature_scale:
        state["center_feature_scale_proj_weight"] = f(G, C, scale=0.2)
        state["center_feature_scale_proj_bias"] = f(G, scale=0.2)
    x = f(m.N, m.H, m.W, C)
    grad = f(m.N, m.H, m.W, C)
    return state, x, grad


MODULE_CASES = [
    ModuleCase("module_c64_g4", seed=0),
    ModuleCase("module_c64_g4_cfs", center_feature_scale=True, seed=1),
    ModuleCase("module_c32_g2_k5_sig2", channels=32, group=2, kernel_size=5, pad=2,
               offset_scale=2.0, H=9, W=11, seed=2),
]
