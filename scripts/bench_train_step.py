"""SURVEY 8f rank 4 numbers: (1) ModelEMA.update, reference loop vs FusedModelEMA on a YOLOv5l-sized tensor set;
(2) DCNv3 layer forward + backward (cfg2 shape), eager vs CUDA-graphed (graph_block).  python scripts/bench_train_step.py"""
import sys, math, time
from copy import deepcopy
import torch
from torch import nn
sys.path.insert(0, '.')
from yolo_somi_b200.train_step import FusedModelEMA, graph_block
from yolo_somi_b200.ops_dcnv3.modules import DCNv3

def stack(n=120, c=256):      # ~ 120 x (conv + bn): 720 state_dict entries, 46 M parameters (YOLOv5l scale)
    layers = []
    for i in range(n):
        layers += [nn.Conv2d(c, c, 3 if i % 3 == 0 else 1, bias=False), nn.BatchNorm2d(c), nn.SiLU()]
    return nn.Sequential(*layers)

model = stack().cuda()
class RefEMA:
    def __init__(self, m): self.ema = deepcopy(m).eval(); self.updates = 0; self.decay = lambda x: 0.9999 * (1 - math.exp(-x / 2000))
    def update(self, model):
        with torch.no_grad():
            self.updates += 1; d = self.decay(self.updates); msd = model.state_dict()
            for k, v in self.ema.state_dict().items():
                if v.dtype.is_floating_point:
                    v *= d; v += (1 - d) * msd[k].detach()
for name, ema in (("reference loop", RefEMA(model)), ("FusedModelEMA", FusedModelEMA(model))):
    for _ in range(3): ema.update(model)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): ema.update(model)
    torch.cuda.synchronize()
    print(f"EMA update, {len(model.state_dict())} tensors, {sum(p.numel() for p in model.parameters())/1e6:.0f} M params: {name:15s} {(time.perf_counter()-t0)/20*1e3:.2f} ms per step", flush=True)

torch.manual_seed(0)
layer = DCNv3(channels=256, group=16).cuda().to(torch.bfloat16)
with torch.no_grad():
    layer.offset.weight.normal_(0, 0.02); layer.mask.weight.normal_(0, 0.02)
x = torch.randn(16, 80, 80, 256, device="cuda", dtype=torch.bfloat16)
graphed = graph_block(layer, (x.clone().requires_grad_(True),))
eager = deepcopy(layer)
gy = torch.randn(16, 80, 80, 256, device="cuda", dtype=torch.bfloat16)
for name, fn, mod in (("eager", eager, eager), ("CUDA graphs", graphed, layer)):
    xs = x.clone().requires_grad_(True)
    for _ in range(5):
        y = fn(xs); torch.autograd.grad(y, [xs] + list(mod.parameters()), gy, allow_unused=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        y = fn(xs); torch.autograd.grad(y, [xs] + list(mod.parameters()), gy, allow_unused=True)
    e1.record(); torch.cuda.synchronize()
    print(f"DCNv3 layer fwd+bwd, N=16 80x80 C=256 G=16 bf16: {name:12s} {e0.elapsed_time(e1)/20:.3f} ms", flush=True)
