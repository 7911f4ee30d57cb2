"""Launch-bound regime: the core on a small feature map (the 20 x 20 head stage of YOLOv5l-DCNv3 at 640 px: N = 16,
C = 512, G = 32) -- GPU time per forward / backward call from CUDA events over back-to-back calls, and host time per
call (enqueue only)."""
import sys, time
import torch
sys.path.insert(0, '.')
import DCNv3
def run(N, H, W, G, gc, iters=200):
    geom = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    g = torch.Generator().manual_seed(1)
    v = torch.randn(N, H, W, G * gc, generator=g); o = torch.randn(N, H, W, G * 18, generator=g)
    m = torch.softmax(torch.randn(N, H, W, G, 9, generator=g), -1).reshape(N, H, W, -1); go = torch.randn(N, H, W, G * gc, generator=g)
    dv, do_, dm, dg = (t.to(torch.bfloat16).cuda() for t in (v, o, m, go))
    for _ in range(10):
        DCNv3.dcnv3_forward(dv, do_, dm, *geom, 256); DCNv3.dcnv3_backward(dv, do_, dm, *geom, dg, 256)
    torch.cuda.synchronize()
    res = {}
    for name, fn in (("fwd", lambda: DCNv3.dcnv3_forward(dv, do_, dm, *geom, 256)), ("bwd", lambda: DCNv3.dcnv3_backward(dv, do_, dm, *geom, dg, 256))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record()
        for _ in range(iters): fn()
        e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
        res[name] = (e0.elapsed_time(e1) * 1e3 / iters, (t1 - t0) * 1e6 / iters)
    print("N %d %dx%d C %d G %d: fwd %.1f us GPU / %.1f us host enqueue, bwd %.1f us GPU / %.1f us host enqueue" %
          (N, H, W, G * gc, G, res["fwd"][0], res["fwd"][1], res["bwd"][0], res["bwd"][1]))
for shape in ((16, 20, 20, 32, 16), (16, 40, 40, 16, 16), (16, 80, 80, 8, 16), (1, 20, 20, 32, 16)):
    run(*shape)
