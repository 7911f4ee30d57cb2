"""Top kernels of YOLOv5l-DCNv3 inference (batch 32, fp16, BatchNorms folded), per batch.  python scripts/infer_prof.py"""
import sys, collections, copy
import torch
sys.path.insert(0, '.')
from torch.profiler import profile, ProfilerActivity
from yolo_somi_b200.yolov5l_dcnv3 import YOLOv5lDCNv3
from yolo_somi_b200.hosting import fuse_for_inference
dev = torch.device('cuda')
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
model = fuse_for_inference(YOLOv5lDCNv3(nc=10).to(dev).to(memory_format=torch.channels_last), half=True)
x = torch.rand(32, 3, 640, 640, device=dev).to(memory_format=torch.channels_last).half()
def one():
    with torch.no_grad():
        return model(x)
for _ in range(3): one()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): one()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if str(getattr(e, "device_type", "")).endswith("CUDA") and e.device_time_total > 0:
        a = agg[e.name[:120]]; a[0] += e.device_time_total; a[1] += 1
tot = sum(v[0] for v in agg.values())
print("GPU time per batch %.2f ms, %d kernels" % (tot / 3e3, sum(v[1] for v in agg.values()) / 3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:32]:
    print("%8.1f us %5.1f %% %4d x  %s" % (v[0] / 3, 100 * v[0] / tot, v[1] / 3, k))
