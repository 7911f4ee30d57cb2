"""Replay stress of the captured forward + backward (the scenario of test_forward_backward_under_cuda_graph_capture):
reports which tensor differs from the eager result, and by how much."""
import sys
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tests/golden')
import DCNv3
import cases
c = cases.Case("graph", N=3, H=26, W=35, G=8, gc=16, seed=411)
v, o, m, g = (torch.as_tensor(a).to(device="cuda", dtype=torch.bfloat16) for a in cases.make_inputs(c))
want = [DCNv3.dcnv3_forward(v, o, m, *c.geom, 256)] + DCNv3.dcnv3_backward(v, o, m, *c.geom, g, 256)
torch.cuda.synchronize()
names = ["out", "grad_value", "grad_offset", "grad_mask"]
# eager repeat
bad = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 200):
    got = [DCNv3.dcnv3_forward(v, o, m, *c.geom, 256)] + DCNv3.dcnv3_backward(v, o, m, *c.geom, g, 256)
    torch.cuda.synchronize()
    for nm, a, w in zip(names, got, want):
        d = float((a.float() - w.float()).abs().max())
        if d > (2e-2 * float(w.float().pow(2).mean().sqrt()) if nm == "grad_value" else 0.0):
            bad += 1; print("eager it", it, nm, "max diff", d, "n diff", int((a != w).sum()))
print("eager mismatches:", bad)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    out = DCNv3.dcnv3_forward(v, o, m, *c.geom, 256)
    grads = DCNv3.dcnv3_backward(v, o, m, *c.geom, g, 256)
bad = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 200):
    for t in [out] + list(grads):
        t.fill_(7.0)
    graph.replay()
    torch.cuda.synchronize()
    for nm, a, w in zip(names, [out] + list(grads), want):
        d = float((a.float() - w.float()).abs().max())
        if d > (2e-2 * float(w.float().pow(2).mean().sqrt()) if nm == "grad_value" else 0.0):
            bad += 1
            idx = (a != w).nonzero()
            print("replay it", it, nm, "max diff", d, "n diff", len(idx), "first", idx[:3].tolist(), "vals", a[tuple(idx[0])].item() if len(idx) else None)
print("replay mismatches:", bad)
