"""Fused dwconv + LayerNorm + GELU (csrc/dcnv3_dwconv.cu) vs the reference's unfused sequence,
BASELINE configs[1] shape (N=16, 80x80, C=256, bf16); CUDA-graph replays, rotating inputs."""
import json, sys
import torch
sys.path.insert(0, '.')
from yolo_somi_b200.ops_dcnv3.functions import dwconv_ln_gelu as dlg

N, H, W, C, k, dt = 16, 80, 80, 256, 3, torch.bfloat16
torch.manual_seed(0)
xs = [torch.randn(N, H, W, C, device='cuda', dtype=dt) for _ in range(4)]
w = (torch.randn(C, 1, k, k, device='cuda') / k).to(dt); b = torch.randn(C, device='cuda').to(dt)
gamma = torch.ones(C, device='cuda', dtype=dt); beta = torch.zeros(C, device='cuda', dtype=dt)

def fused(x):
    with torch.no_grad():
        return dlg.DwConvLnGelu.apply(x, w, b, gamma, beta, 1e-6, dt)

def eager(x):
    with torch.no_grad():
        return dlg._unfused(x, w, b, gamma, beta, 1e-6)

def graph_us(fn, reps=8, replays=6):
    for i in range(3): fn(xs[i % 4])
    torch.cuda.synchronize()
    g, s = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            for i in range(reps): fn(xs[i % 4])
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * replays)

t_f, t_e = graph_us(fused), graph_us(eager)
by = 2 * 2 * N * H * W * C
print(json.dumps({"shape": {"N": N, "H": H, "W": W, "C": C, "k": k, "dtype": "bf16"}, "fused_us": t_f, "eager_us": t_e,
                  "algorithmic_bytes": by, "fused_gbs": by / t_f / 1e3,
                  "note": "fused time includes the wrapper's small parameter casts"}))
