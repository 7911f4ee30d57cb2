"""grad_value with the accumulator resident in tensor memory (default) against the plane form (DCNV3_VALUE=plane) on
shapes with few patch rows per CTA: backward per call, CUDA events.  python scripts/vres_vs_plane.py"""
import os, sys
import torch
sys.path.insert(0, '.')
import DCNv3

def bwd_us(N, H, W, G, iters=30):
    geom = (3, 3, 1, 1, 1, 1, 1, 1, G, 16, 1.0)
    g = torch.Generator().manual_seed(1)
    v = torch.randn(N, H, W, G * 16, generator=g); o = torch.randn(N, H, W, G * 18, generator=g)
    m = torch.softmax(torch.randn(N, H, W, G, 9, generator=g), -1).reshape(N, H, W, -1); go = torch.randn(N, H, W, G * 16, generator=g)
    t = [x.bfloat16().cuda() for x in (v, o, m, go)]
    out = {}
    for mode in ("default", "plane", "default", "plane"):
        if mode == "plane": os.environ["DCNV3_VALUE"] = "plane"
        else: os.environ.pop("DCNV3_VALUE", None)
        for _ in range(3): DCNv3.dcnv3_backward(*t[:3], *geom, t[3], 256)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): DCNv3.dcnv3_backward(*t[:3], *geom, t[3], 256)
        e1.record(); torch.cuda.synchronize()
        out.setdefault(mode, []).append(e0.elapsed_time(e1) * 1e3 / iters)
    os.environ.pop("DCNV3_VALUE", None)
    rows = N * G * ((H + 7) // 8)
    print("N %3d %3dx%-3d G %2d  rows/148 %5.1f   default %s us   plane %s us" % (N, H, W, G, rows / 148.0,
          " / ".join("%.1f" % x for x in out["default"]), " / ".join("%.1f" % x for x in out["plane"])), flush=True)

shapes = [(1, 80, 80, 16), (2, 80, 80, 16), (4, 80, 80, 16), (8, 80, 80, 16), (16, 40, 40, 16), (16, 20, 20, 32), (16, 80, 80, 8), (2, 40, 40, 16), (1, 152, 152, 16)]
if len(sys.argv) > 1 and sys.argv[1] == "large":
    shapes = [(128, 20, 20, 32), (32, 20, 20, 32), (128, 40, 40, 16), (64, 40, 40, 16), (6, 80, 80, 16), (5, 80, 80, 16), (12, 80, 80, 16), (32, 80, 80, 8)]
for shape in shapes:
    bwd_us(*shape)
