"""Cycle stamps of CTA 0 of the resident-accumulator value kernel (DCNV3_VRES_DIAG=1): per patch, where the builder
group, the refill / input warp and the product warp spend their time."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, '.')
os.environ['DCNV3_VRES_DIAG'] = '1'
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
buf = np.zeros((256, 8), dtype=np.int64)
rc = lib.dcnv3_vres_debug_read(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0, rc
n = int((buf[:, 7] > 0).sum())
t0 = buf[0, 0]
print("patch | builder: start +wait_inputs +wait_tile +build | loader: refill issued | products: start +wait_full +issue   (cycles from the first stamp)")
for p in range(min(n, 40)):
    r = buf[p] - t0
    print(f"{p:4d} | b {r[0]:7d} +{r[3]-r[0]:6d} +{r[1]-r[3]:6d} +{r[2]-r[1]:6d} | l {r[4]:7d} | m {r[5]:7d} +{r[6]-r[5]:6d} +{r[7]-r[6]:5d}")
r = buf[20:n]
print("means over patches 20..%d:" % n)
print("  builder wait_inputs %.0f  wait_tile %.0f  build %.0f" % ((r[:, 3]-r[:, 0]).mean(), (r[:, 1]-r[:, 3]).mean(), (r[:, 2]-r[:, 1]).mean()))
print("  products wait_full %.0f  issue %.0f" % ((r[:, 6]-r[:, 5]).mean(), (r[:, 7]-r[:, 6]).mean()))
print("  patch-to-patch (product issue) %.0f cycles" % np.diff(buf[20:n, 7]).mean())
print("  builder group: end of build(p) -> start of patch p+4: %.0f" % (buf[24:n, 0] - buf[20:n-4, 2]).mean())
print("  refill issued(p) -> tile ready seen by builder: %.0f" % (buf[20:n, 1] - buf[20:n, 4]).mean())
print("  products done issuing(p) -> refill issued for p+6: %.0f" % (buf[26:n, 4] - buf[20:n-6, 7]).mean())
