import sys, torch
sys.path.insert(0, '.')
from torch.profiler import profile, ProfilerActivity
from yolo_somi_b200.ops_dcnv3.modules import DCNv3 as Layer
dev, dt = torch.device('cuda'), torch.bfloat16
torch.manual_seed(0)
layer = Layer(channels=256, group=16).to(dev).to(dt)
with torch.no_grad():
    layer.offset.weight.normal_(0, 0.02); layer.mask.weight.normal_(0, 0.1)
x = torch.randn(16, 80, 80, 256, device=dev, dtype=dt, requires_grad=True)
go = torch.randn(16, 80, 80, 256, device=dev, dtype=dt)
for _ in range(3): layer(x).backward(go)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): layer(x).backward(go)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=70))
