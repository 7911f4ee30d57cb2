import sys, time
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tests/golden')
from oracle import dcnv3_oracle as orc
import DCNv3
for G in (8, 16, 32):
    gc = 256 // G
    N, H, W = 1, 192, 192
    geom = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
    g = torch.Generator().manual_seed(G)
    v = torch.randn(N, H, W, 256, generator=g); o = torch.randn(N, H, W, G * 18, generator=g)
    m = torch.softmax(torch.randn(N, H, W, G, 9, generator=g), -1).reshape(N, H, W, -1); go = torch.randn(N, H, W, 256, generator=g)
    arrs = [t.to(torch.bfloat16) for t in (v, o, m, go)]
    dv, do_, dm, dg = (t.cuda() for t in arrs)
    out = DCNv3.dcnv3_forward(dv, do_, dm, *geom, 256)
    grads = DCNv3.dcnv3_backward(dv, do_, dm, *geom, dg, 256)
    torch.cuda.synchronize()
    # timing
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for _ in range(3):
        DCNv3.dcnv3_forward(dv, do_, dm, *geom, 256); DCNv3.dcnv3_backward(dv, do_, dm, *geom, dg, 256)
    e[0].record()
    for _ in range(10): DCNv3.dcnv3_forward(dv, do_, dm, *geom, 256)
    e[1].record()
    for _ in range(10): DCNv3.dcnv3_backward(dv, do_, dm, *geom, dg, 256)
    e[2].record(); torch.cuda.synchronize()
    a64 = [t.double().numpy() for t in arrs]
    want = (orc.direct_forward(*a64[:3], *geom), *orc.direct_backward(*a64, *geom))
    res = []
    for nm, a, w in zip(("out", "gv", "go", "gm"), (out, *grads), want):
        a = a.double().cpu().numpy(); rms = float(np.sqrt(np.mean(w ** 2)))
        res.append((nm, float(np.mean(np.abs(a - w) > 1e-2 * np.abs(w) + 1e-2 * rms))))
    pts = N * H * W * G * 9
    print(f"G={G} gc={gc}: fwd {e[0].elapsed_time(e[1])*100:.1f} us bwd {e[1].elapsed_time(e[2])*100:.1f} us  pts {pts/1e6:.1f}M  bad-frac {res}", flush=True)
