import sys, time, torch
sys.path.insert(0, '.')
import bench
from yolo_somi_b200.host_pipeline import DCNv3HostPipeline
dev = torch.device('cuda', 0); torch.cuda.set_device(dev)
C = bench.CFG
ins = [t.cpu().pin_memory() for t in bench.make_inputs(C['N'], dev, torch.bfloat16, 0)]
for chunk in (1, 2, 4, 8, 16):
    pipe = DCNv3HostPipeline(C['H'], C['W'], C['G'], C['C']//C['G'], dtype=torch.bfloat16, chunk_images=chunk, device=dev)
    sv, so, sm, sy = pipe.shapes(C['N'])
    outs = [torch.empty(s, dtype=torch.bfloat16).pin_memory() for s in (sy, sv, so, sm)]
    for _ in range(2): pipe.run(*ins, *outs)
    pipe.sync()
    t0 = time.perf_counter()
    for _ in range(20): pipe.run(*ins, *outs)
    pipe.sync()
    dt = (time.perf_counter() - t0) / 20
    print(f'chunk {chunk:2d}: {dt*1e3:.2f} ms/step  {193.3312/dt/1e3:.1f} GB/s each way', flush=True)
    pipe.close()
# plain copies for reference
d = [t.to(dev) for t in ins]
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    for h, x in zip(ins, d): x.copy_(h, non_blocking=True)
torch.cuda.synchronize(); print('H2D only GB/s', 10*193.3312/ (time.perf_counter()-t0)/1e3)
