"""Top kernels of the YOLOv5l-DCNv3 training step (fp16 AMP, channels_last), per step, and the same for one hosted
DCNv3_YOLO block under autocast -- which elementwise / copy kernels belong to this library's layer and which to the
host zoo's Conv / C3 / Concat.  python scripts/train_prof.py [batch]"""
import sys, collections
import torch
sys.path.insert(0, '.')
from torch.profiler import profile, ProfilerActivity
from yolo_somi_b200.train_step import TrainStep, make_optimizer, synthetic_batch
from yolo_somi_b200.yolov5l_dcnv3 import YOLOv5lDCNv3
from yolo_somi_b200.hosting import DCNv3_YOLO

def table(prof, steps, top=28):
    agg = collections.defaultdict(lambda: [0.0, 0])
    for e in prof.events():
        if getattr(e, "device_type", None) is not None and str(e.device_type).endswith("CUDA") and e.device_time_total > 0:
            a = agg[e.name[:110]]; a[0] += e.device_time_total; a[1] += 1
    tot = sum(v[0] for v in agg.values())
    print("total GPU time per step %.2f ms, %d kernels per step" % (tot / steps / 1e3, sum(v[1] for v in agg.values()) / steps))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%8.1f us %5.1f %% %5d x  %s" % (v[0] / steps, 100 * v[0] / tot, v[1] / steps, k))

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device('cuda')
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
model = YOLOv5lDCNv3(nc=10).to(dev).to(memory_format=torch.channels_last)
ts = TrainStep(model, nc=10, optimizer=make_optimizer(model), autocast_dtype=torch.float16)
imgs, targets = synthetic_batch(B, 640, device="cuda")
imgs = imgs.to(memory_format=torch.channels_last)
import os
import yolo_somi_b200.train_step as _ts
_k = float(os.environ.get("TRAIN_LOSS_SCALE", 1))
if _k != 1:      # what `loss * WORLD_SIZE` does to the updates at N ranks, on one GPU
    _orig = _ts.yolo_surrogate_loss
    _ts.yolo_surrogate_loss = lambda *a: _orig(*a) * _k
for _ in range(int(os.environ.get("TRAIN_WARM", 3))): ts.step(imgs, targets)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2): ts.step(imgs, targets)
    torch.cuda.synchronize()
print("== training step, batch %d" % B); table(prof, 2, int(os.environ.get("TOP", 28)))
if os.environ.get("STEP_ONLY"): sys.exit(0)
del ts, model
blk = DCNv3_YOLO(256, 256, 3).to(dev).to(memory_format=torch.channels_last)
x = torch.randn(B, 256, 40, 40, device=dev).to(memory_format=torch.channels_last).requires_grad_(True)
def one():
    with torch.autocast("cuda", dtype=torch.float16):
        y = blk(x)
    y.float().square().mean().backward()
for _ in range(3): one()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): one()
    torch.cuda.synchronize()
print("== one DCNv3_YOLO block (C = 256, 40 x 40, batch %d) under fp16 autocast" % B); table(prof, 3, 40)
