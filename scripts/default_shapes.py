"""Timing rows for shapes outside (or at the edge of) the fastest kernels' gates: the reference layer's constructor
defaults (C = 64, G = 4: modules/dcnv3.py:223-237), the reference test script's offset_scale = 2 (test.py:19-30), stride 2,
5 x 5 kernels, fp32 I/O, and BASELINE configs[4] (192 x 192, C = 256, G = 8 / 16 / 32).  Forward / backward per call from
CUDA events over back-to-back calls; which kernels ran is in the launch list (scripts/gpu_launches.sh)."""
import sys
import torch
sys.path.insert(0, '.')
import DCNv3

def run(name, N, H, W, G, gc, k=3, s=1, p=None, d=1, sigma=1.0, dtype=torch.bfloat16, iters=30):
    p = (d * (k - 1)) // 2 if p is None else p
    Ho = (H + 2 * p - (d * (k - 1) + 1)) // s + 1; Wo = (W + 2 * p - (d * (k - 1) + 1)) // s + 1
    geom = (k, k, s, s, p, p, d, d, G, gc, sigma)
    g = torch.Generator().manual_seed(1)
    v = torch.randn(N, H, W, G * gc, generator=g); o = torch.randn(N, Ho, Wo, G * k * k * 2, generator=g)
    m = torch.softmax(torch.randn(N, Ho, Wo, G, k * k, generator=g), -1).reshape(N, Ho, Wo, -1); go = torch.randn(N, Ho, Wo, G * gc, generator=g)
    dv, do_, dm, dg = (t.to(dtype).cuda() for t in (v, o, m, go))
    for _ in range(3):
        DCNv3.dcnv3_forward(dv, do_, dm, *geom, 256); DCNv3.dcnv3_backward(dv, do_, dm, *geom, dg, 256)
    torch.cuda.synchronize()
    t = {}
    for nm, fn in (("fwd", lambda: DCNv3.dcnv3_forward(dv, do_, dm, *geom, 256)), ("bwd", lambda: DCNv3.dcnv3_backward(dv, do_, dm, *geom, dg, 256))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): fn()
        e1.record(); torch.cuda.synchronize()
        t[nm] = e0.elapsed_time(e1) * 1e3 / iters
    pts = N * Ho * Wo * G * k * k
    print("%-34s N %2d %3dx%-3d C %4d G %2d k %d s %d sigma %.1f %-8s fwd %7.1f us  bwd %7.1f us  (%.1f M points, %.2f / %.2f ns per k points)" %
          (name, N, H, W, G * gc, G, k, s, sigma, str(dtype).split('.')[-1], t["fwd"], t["bwd"], pts / 1e6, t["fwd"] * 1e3 / (pts / 1e3), t["bwd"] * 1e3 / (pts / 1e3)))

run("cfg2 (reference point)", 16, 80, 80, 16, 16)
run("layer default C=64 G=4", 16, 80, 80, 4, 16)
run("test.py offset_scale=2", 16, 80, 80, 16, 16, sigma=2.0)
run("stride 2", 16, 80, 80, 16, 16, s=2)
run("5x5 kernel", 16, 80, 80, 16, 16, k=5)
run("dilation 2", 16, 80, 80, 16, 16, d=2)
run("fp32 I/O", 16, 80, 80, 16, 16, dtype=torch.float32)
run("fp16 I/O", 16, 80, 80, 16, 16, dtype=torch.float16)
run("fp64 I/O (correctness path)", 4, 80, 80, 16, 16, dtype=torch.float64, iters=5)
run("map 240 px wide (ring of two)", 2, 64, 240, 16, 16)
run("map 248 px wide (plane form)", 2, 64, 248, 16, 16)
run("cfg1 (reference CPU case) fp32", 2, 32, 32, 4, 16, dtype=torch.float32)
for G in (8, 16, 32):
    run("cfg5 192x192 G=%d" % G, 1, 192, 192, G, 256 // G)
    run("cfg5 192x192 G=%d N=4" % G, 4, 192, 192, G, 256 // G)
