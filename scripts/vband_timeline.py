"""Cycle stamps of CTA 0 of the dense-band value kernel (experiments build, DCNV3_VBAND_DIAG |= 64): where a link of the
product -> drain -> zero -> product chain and a builder half-patch spend their time."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, '.')
os.environ['DCNV3_VALUE'] = 'band'
os.environ['DCNV3_VBAND_DIAG'] = str(64 | int(os.environ.get('EXTRA_DIAG', '0')))
import DCNv3
from yolo_somi_b200 import _native
lib = _native.load()
N, H, W, G, gc = 16, 80, 80, 16, 16
geom = (3, 3, 1, 1, 1, 1, 1, 1, G, gc, 1.0)
g = torch.Generator().manual_seed(1)
v = torch.randn(N, H, W, G * gc, generator=g); o = torch.randn(N, H, W, G * 18, generator=g)
m = torch.softmax(torch.randn(N, H, W, G, 9, generator=g), -1).reshape(N, H, W, -1); go = torch.randn(N, H, W, G * gc, generator=g)
dv, do_, dm, dg = (t.to(torch.bfloat16).cuda() for t in (v, o, m, go))
for _ in range(3):
    DCNv3.dcnv3_backward(dv, do_, dm, *geom, dg, 256)
torch.cuda.synchronize()
buf = np.zeros((2, 256, 8), dtype=np.int64)
rc = lib.dcnv3_vband_debug_read(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0, rc
d, b = buf[0], buf[1]
ok = d[:, 2] > 0
links = np.nonzero(ok)[0]
t0 = d[links[0], 0]
print("drain side (cycles): link  wait_token  wait_commit  ld  zero  issue_next  reduce | link-to-link")
prev = None
rows = []
for l in links[:80]:
    r = d[l]
    rows.append((r[1]-r[0], r[2]-r[1], r[3]-r[2], r[4]-r[3], r[5]-r[4], r[6]-r[5], (r[2]-prev) if prev else 0))
    prev = r[2]
rows = np.array(rows)
for l, r in zip(links[:24], rows[:24]):
    print(f"  {l:3d}  " + "  ".join(f"{x:7d}" for x in r))
print("mean over links 8..80:", np.round(rows[8:].mean(0)).astype(int))
okb = b[:, 4] > 0
hp = np.nonzero(okb)[0]
rb = np.array([(b[h, 1]-b[h, 0], b[h, 2]-b[h, 1], b[h, 3]-b[h, 2], b[h, 4]-b[h, 3], (b[h, 0]-b[h-1, 0]) if h else 0) for h in hp[:60]])
print("builder warp 0 (cycles): wait_a_free  wait_inputs  nine_points  stores+far+arrive | hp-to-hp")
print("mean over hp 4..60:", np.round(rb[4:].mean(0)).astype(int))
