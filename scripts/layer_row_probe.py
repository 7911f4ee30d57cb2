"""Is the layer row GPU-bound?  GPU time (events) and host enqueue time per forward + backward of the cfg2 layer, with the
bias-gradient column sums as fp32-out / 16-bit-out ones-row GEMMs.  python scripts/layer_row_probe.py"""
import sys, time, torch
sys.path.insert(0, '.')
from yolo_somi_b200.ops_dcnv3.modules import DCNv3 as Layer
from yolo_somi_b200.ops_dcnv3.functions import offset_mask_proj as omp
dev, dt = torch.device('cuda'), torch.bfloat16
torch.manual_seed(0)
layer = Layer(channels=256, group=16).to(dev).to(dt)
with torch.no_grad():
    layer.offset.weight.normal_(0, 0.02); layer.mask.weight.normal_(0, 0.1)
xs = [torch.randn(16, 80, 80, 256, device=dev, dtype=dt, requires_grad=True) for _ in range(2)]
go = torch.randn(16, 80, 80, 256, device=dev, dtype=dt)
def timed(n=40):
    for i in range(6): layer(xs[i % 2]).backward(go)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(n): layer(xs[i % 2]).backward(go)
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (t1 - t0) * 1e3 / n
print("fp32-out column sums: gpu %.3f ms, host enqueue %.3f ms" % timed())
orig = omp.column_sums
def cs16(g2):
    return (torch.ones(1, g2.shape[0], dtype=g2.dtype, device=g2.device) @ g2).reshape(-1)
omp.column_sums = cs16
print("16-bit-out column sums: gpu %.3f ms, host enqueue %.3f ms" % timed())
omp.column_sums = orig
print("fp32-out again: gpu %.3f ms, host enqueue %.3f ms" % timed())
g = torch.randn(102400, 256, device=dev, dtype=dt)
for name, fn in (("fp32-out", orig), ("16-bit-out", cs16)):
    for _ in range(3): fn(g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn(g)
    e1.record(); torch.cuda.synchronize()
    print("column sums of [102400, 256] %s: %.1f us" % (name, e0.elapsed_time(e1) * 50))
