import torch, time
torch.backends.cudnn.benchmark=True
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/n
N,C,H,W=128,256,80,80
for dt in (torch.bfloat16, torch.float16, torch.float32):
  for cl in (True, False):
    for cudnn in (True, False):
        torch.backends.cudnn.enabled = cudnn
        x=torch.randn(N,C,H,W,device='cuda',dtype=dt)
        if cl: x=x.to(memory_format=torch.channels_last)
        x.requires_grad_(True)
        bn=torch.nn.BatchNorm2d(C).cuda().train()
        g=torch.randn_like(x)
        def f():
            y=bn(x); y.backward(g)
        ms=t(f)
        gb = x.numel()*x.element_size()/1e9
        print(f"{str(dt)[6:]:9s} channels_last={cl!s:5s} cudnn={cudnn!s:5s}: {ms:.3f} ms  ({gb:.2f} GB tensor -> {gb*5/ms*1e3:.0f} GB/s at 5 passes)")
torch.backends.cudnn.enabled=True
# SiLU
x=torch.randn(N,C,H,W,device='cuda',dtype=torch.bfloat16).to(memory_format=torch.channels_last).requires_grad_(True)
act=torch.nn.SiLU()
g=torch.randn_like(x)
print("silu fwd+bwd", t(lambda: act(x).backward(g)))
