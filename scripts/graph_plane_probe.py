"""The plane form of the backward (small shapes since the size heuristic) under CUDA-graph capture: replay == eager."""
import sys, torch
sys.path.insert(0, '.')
import DCNv3
N, H, W, G = 16, 20, 20, 32
geom = (3, 3, 1, 1, 1, 1, 1, 1, G, 16, 1.0)
g = torch.Generator().manual_seed(2)
t = [x.bfloat16().cuda() for x in (torch.randn(N, H, W, G * 16, generator=g), torch.randn(N, H, W, G * 18, generator=g),
     torch.softmax(torch.randn(N, H, W, G, 9, generator=g), -1).reshape(N, H, W, -1), torch.randn(N, H, W, G * 16, generator=g))]
eager = DCNv3.dcnv3_backward(*t[:3], *geom, t[3], 256)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    DCNv3.dcnv3_backward(*t[:3], *geom, t[3], 256)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s):
        out = DCNv3.dcnv3_backward(*t[:3], *geom, t[3], 256)
gr.replay(); torch.cuda.synchronize()
print("captured replay == eager:", [bool(torch.equal(a, b)) for a, b in zip(out[1:], eager[1:])],
      "grad_value max diff", float((out[0].float() - eager[0].float()).abs().max()))
