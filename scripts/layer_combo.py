import os, sys, torch
sys.path.insert(0, '.')
from yolo_somi_b200.ops_dcnv3.modules import DCNv3 as Layer
dev, dt = torch.device('cuda'), torch.bfloat16
torch.manual_seed(0)
layer = Layer(channels=256, group=16).to(dev).to(dt)
with torch.no_grad():
    layer.offset.weight.normal_(0, 0.02); layer.mask.weight.normal_(0, 0.1)
xs = [torch.randn(16, 80, 80, 256, device=dev, dtype=dt, requires_grad=True) for _ in range(2)]
go = torch.randn(16, 80, 80, 256, device=dev, dtype=dt)
def timed(train, n=10):
    def step(i):
        if train: layer(xs[i % 2]).backward(go)
        else:
            with torch.no_grad(): layer(xs[i % 2])
    for i in range(3): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): step(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for proj in ("1", "0"):
    for dw in ("1", "0"):
        os.environ["DCNV3_FUSED_PROJ"] = proj; os.environ["DCNV3_FUSED_DWCONV"] = dw
        print(f"proj={proj} dwconv={dw}: fwd {timed(False):.3f} ms  fwd+bwd {timed(True):.3f} ms", flush=True)
