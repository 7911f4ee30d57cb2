"""Fused offset/mask projection (tcgen05) vs the layer's two linears + softmax on cuBLAS/eager,
BASELINE configs[1] shape: M = 16*80*80 rows, C = 256, G = 16.  Prints one JSON object."""
import json, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from yolo_somi_b200.ops_dcnv3.functions import offset_mask_proj as omp

M, C, G, dt = 16 * 80 * 80, 256, 16, torch.bfloat16
torch.manual_seed(0)
xs = [torch.randn(M, C, device='cuda', dtype=dt) for _ in range(4)]      # rotate: > L2 between uses
w_off = (torch.randn(2 * G * 9, C, device='cuda') / 16).to(dt); b_off = torch.randn(2 * G * 9, device='cuda').to(dt)
w_msk = (torch.randn(G * 9, C, device='cuda') / 8).to(dt); b_msk = torch.randn(G * 9, device='cuda').to(dt)

def fused(x):
    with torch.no_grad():
        return omp.OffsetMaskProj.apply(x, w_off, b_off, w_msk, b_msk, G, dt)

def eager(x):   # modules/dcnv3.py:330-334
    with torch.no_grad():
        off = F.linear(x, w_off, b_off)
        msk = F.softmax(F.linear(x, w_msk, b_msk).reshape(M, G, -1).float(), -1).reshape(M, -1).to(dt)
        return off, msk

def timeit(fn, n=30):
    for i in range(5): fn(xs[i % 4])
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    e[0].record()
    for i in range(n):
        fn(xs[i % 4]); e[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(e[i].elapsed_time(e[i + 1]) for i in range(n))
    return ts[n // 2] * 1e3

t_f, t_e = timeit(fused), timeit(eager)
by = 2 * (M * C + M * 3 * G * 9) + 2 * 3 * G * 9 * C
fl = 2.0 * M * C * 3 * G * 9
print(json.dumps({"shape": {"M": M, "C": C, "G": G, "dtype": "bf16"}, "fused_us": t_f, "eager_us": t_e,
                  "algorithmic_bytes": by, "fused_gbs": by / t_f / 1e3, "fused_tflops": fl / t_f / 1e6,
                  "note": "fused time includes the per-call weight packing (torch.cat + cast) of the Python wrapper"}))
