"""Which Python line of this library launches which copy / cast / reduction kernel inside one hosted DCNv3 block
(Bottleneck_DCNv3: 1x1 Conv -> DCNv3_YOLO -> residual) under fp16 autocast, forward + backward, with activations in
the layout and dtype the host model hands over (fp16, channels_last).  python scripts/block_copies.py [batch] [C] [HW]"""
import sys, collections
import torch
sys.path.insert(0, '.')
from torch.profiler import profile, ProfilerActivity
from yolo_somi_b200.hosting import Bottleneck_DCNv3

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
C = int(sys.argv[2]) if len(sys.argv) > 2 else 256
HW = int(sys.argv[3]) if len(sys.argv) > 3 else 40
dev = torch.device('cuda')
torch.manual_seed(0)
blk = Bottleneck_DCNv3(C, C, True, e=1.0).to(dev).to(memory_format=torch.channels_last)
x = torch.randn(B, C, HW, HW, device=dev, dtype=torch.float16).to(memory_format=torch.channels_last).requires_grad_(True)
gy = torch.randn(B, C, HW, HW, device=dev, dtype=torch.float16).to(memory_format=torch.channels_last)
def one():
    for p in blk.parameters(): p.grad = None
    with torch.autocast("cuda", dtype=torch.float16):
        y = blk(x)
    y.backward(gy)
for _ in range(3): one()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    one(); torch.cuda.synchronize()
tot = 0.0; rows = collections.defaultdict(lambda: [0.0, 0])
kern = collections.defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if str(getattr(e, "device_type", "")).endswith("CUDA"):
        kern[e.name[:90]][0] += e.device_time_total; kern[e.name[:90]][1] += 1; tot += e.device_time_total
print("GPU time %.1f us, %d kernels" % (tot, sum(v[1] for v in kern.values())))
for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])[:40]:
    print("%8.1f us %3d x  %s" % (v[0], v[1], k))
watch = ("aten::copy_", "aten::sum", "aten::cat", "aten::fill_", "aten::zero_", "aten::add_", "aten::add", "aten::mul", "aten::clone")
print("== elementwise / copy / reduction ops by the innermost frame of this repo")
for e in prof.events():
    if e.name in watch and e.device_time_total > 0 and not str(getattr(e, "device_type", "")).endswith("CUDA"):
        fr = [s for s in (e.stack or []) if "/yolo_somi_b200/" in s or "hosting" in s]
        where = fr[0].split("/yolo_somi_b200/")[-1] if fr else ((e.stack or ["?"])[0][-60:])
        rows[(e.name, where, str(e.input_shapes)[:60])][0] += e.device_time_total; rows[(e.name, where, str(e.input_shapes)[:60])][1] += 1
for k, v in sorted(rows.items(), key=lambda kv: -kv[1][0])[:45]:
    print("%8.1f us %3d x  %-12s %-58s %s" % (v[0], v[1], k[0], k[1][:58], k[2]))
