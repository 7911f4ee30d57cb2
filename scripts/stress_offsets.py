"""cfg2 (N=16, 80x80, C=256, G=16, bf16) with wider offset distributions ('also report a stress run
offset ~ U(-4, 4)'): fwd / bwd CUDA-event times of the default kernels.  python scripts/stress_offsets.py"""
import os, sys
import torch
sys.path.insert(0, '.')
import DCNv3

N, H, W, G, gc = 16, 80, 80, 16, 16
geom = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
g = torch.Generator().manual_seed(0)
v = torch.randn(N, H, W, G * gc, generator=g).bfloat16().cuda()
m = torch.softmax(torch.randn(N, H, W, G, 9, generator=g), -1).reshape(N, H, W, -1).bfloat16().cuda()
go = torch.randn(N, H, W, G * gc, generator=g).bfloat16().cuda()
dists = {"N(0,0.5)": lambda: 0.5 * torch.randn(N, H, W, G * 18, generator=g),
         "N(0,1)": lambda: torch.randn(N, H, W, G * 18, generator=g),
         "N(0,2)": lambda: 2.0 * torch.randn(N, H, W, G * 18, generator=g),
         "U(-4,4)": lambda: 8.0 * torch.rand(N, H, W, G * 18, generator=g) - 4.0}
only = os.environ.get('STRESS_ONLY')   # e.g. 'N(0,2)': one distribution (for an ncu launch list)
for name, mk in dists.items():
    if only and name != only:
        continue
    o = mk().bfloat16().cuda()
    for _ in range(3):
        DCNv3.dcnv3_forward(v, o, m, *geom, 256); DCNv3.dcnv3_backward(v, o, m, *geom, go, 256)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(10): DCNv3.dcnv3_forward(v, o, m, *geom, 256)
    e[1].record()
    for _ in range(10): DCNv3.dcnv3_backward(v, o, m, *geom, go, 256)
    e[2].record(); torch.cuda.synchronize()
    print(f"{os.environ.get('DCNV3_BWD', 'default'):8s} offsets {name:9s}: fwd {e[0].elapsed_time(e[1]) * 100:7.1f} us  bwd {e[1].elapsed_time(e[2]) * 100:7.1f} us", flush=True)
