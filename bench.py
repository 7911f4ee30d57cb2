#!/usr/bin/env python
"""bench.py -- DCNv3 core fwd+bwd throughput on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A *step* is one forward + one backward of the DCNv3 core over one batch of synthetic tensors of
BASELINE.json configs[1]: N=16 per GPU, 80x80, C=256, group=16, 3x3, stride 1, pad 1, bf16
(value~N(0,1), offset~N(0,1) px, mask=softmax(N(0,1)), grad_out~N(0,1), sigma=1).
Unit of work: one sampled point (n,ho,wo,g,p); a step processes N*Ho*Wo*G*9 = 14,745,600 per GPU.

  value      whole-job sampled-points/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the public API with HOST (pinned) buffers: H2D of the four inputs
             and D2H of the four results inside the timed region
  roofline   backward pass (the dominant launches), algorithmic bytes / event time vs measured HBM
  cpu_baseline  the reference's dcnv3_core_pytorch (baseline/_ref, else the oracle's port) on this box's host cores,
             the full workload, a few passes (rank 0, N=1 only)

Multi-GPU: the batch shards over ranks with no data-path collective -> weak scaling.
`--impl reference` times the reference's own CPU path -- the unmodified dcnv3_core_pytorch from the verbatim copy
staged under git-ignored baseline/_ref/ (scripts/stage_reference.py) -- on all 16 images, host cores, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "dcnv3_core_fwd_bwd_sampled_points_per_s"
UNIT = "sampled-pts/s"
CFG = dict(N=16, H=80, W=80, C=256, G=16, K=3, stride=1, pad=1, dil=1, sigma=1.0, dtype="bf16")
WORKLOAD = "DCNv3 core fwd+bwd, N=16/GPU H=W=80 C=256 group=16 k=3 s=1 p=1 bf16 (BASELINE configs[1])"
ROTATE = 3  # input sets cycled so that a step never finds its inputs in the 126 MB L2


def geom():
    gc = CFG["C"] // CFG["G"]
    k, s, p, d = CFG["K"], CFG["stride"], CFG["pad"], CFG["dil"]
    return (k, k, s, s, p, p, d, d, CFG["G"], gc, CFG["sigma"])


def points_per_step(n=None):
    n = CFG["N"] if n is None else n
    return n * CFG["H"] * CFG["W"] * CFG["G"] * CFG["K"] ** 2


def algorithmic_bytes(n=None, s=2):
    """fwd = s(V+O+3Q), bwd = s(2V+O+6Q) with every tensor touched once."""
    n = CFG["N"] if n is None else n
    V = n * CFG["H"] * CFG["W"] * CFG["C"]
    O = V
    Q = points_per_step(n)
    return dict(fwd=s * (V + O + 3 * Q), bwd=s * (2 * V + O + 6 * Q))


def make_inputs(n, device, dtype, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    H, W, C, G, P = CFG["H"], CFG["W"], CFG["C"], CFG["G"], CFG["K"] ** 2
    value = torch.randn(n, H, W, C, generator=g)
    offset = torch.randn(n, H, W, G * P * 2, generator=g)
    mask = torch.softmax(torch.randn(n, H, W, G, P, generator=g), -1).reshape(n, H, W, G * P)
    grad = torch.randn(n, H, W, C, generator=g)
    return tuple(t.to(dtype).to(device) for t in (value, offset, mask, grad))


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    (profiles/traffic.json: {"kernel": ..., "dram_bytes": read+write, "source": file})."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("dram_bytes"), d.get("source")
    return None, None


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every 2 ms from a thread (the
    timed region is a few milliseconds, `nvidia-smi -lms 100` would see it once); nvidia-smi is the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.stop = index, [], None, None, threading.Event()

    def _visible_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v.strip() for v in vis.split(",") if v.strip()]
        if self.index < len(ids) and ids[self.index].isdigit():
            return int(ids[self.index])
        return self.index

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._visible_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _poll(self):
        nv = self.nvml
        bits = (("hw_slowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                ("hw_thermal_slowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                ("sw_thermal_slowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                ("sw_power_cap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)))
        while not self.stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.rows.append([str(mhz), str(self.max_mhz)] + ["Active" if mask & b else "Not Active" for _, b in bits])
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.nvml:
            self.stop.set()
            self.thread.join(timeout=2)
            try:
                self.nvml.nvmlShutdown()
            except Exception:
                pass
        elif self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml, 2 ms poll" if self.nvml else "nvidia-smi -lms 100"}


# ------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_fn():
    """The reference's own CPU path: UNMODIFIED dcnv3_core_pytorch (functions/dcnv3_func.py:147-188) from the verbatim
    copy staged under git-ignored baseline/_ref/ (scripts/stage_reference.py; it travels with the snapshot).  Only if
    that copy is absent: the oracle's restatement of the same function (kind "port")."""
    from baseline import ref_loader
    if ref_loader.available():
        core = ref_loader.core_pytorch()

        def fwd_bwd(v, o, m, g, *geo):
            v = v.clone().requires_grad_(True); o = o.clone().requires_grad_(True); m = m.clone().requires_grad_(True)
            out = core(v, o, m, *geo)
            out.backward(g)
            return out, v.grad, o.grad, m.grad
        return fwd_bwd, "reference", "unmodified dcnv3_core_pytorch from baseline/_ref (grid_sample + autograd)"
    from oracle import dcnv3_oracle as orc
    return orc.gridsample_fwd_bwd, "port", "oracle port of dcnv3_core_pytorch (grid_sample + autograd)"


def cpu_reference_run(sample_n, repeats, warm=1):
    """fwd+bwd of the CPU reference (fp32 arithmetic on the bf16-rounded inputs of the workload) on `sample_n` images;
    returns (list of seconds per pass, threads, kind, description)."""
    fn, kind, desc = cpu_reference_fn()
    torch.set_num_threads(os.cpu_count() or 1)
    v, o, m, g = make_inputs(sample_n, "cpu", torch.bfloat16, seed=1234)
    v, o, m, g = (t.float() for t in (v, o, m, g))     # bf16-rounded values, fp32 arithmetic
    times = []
    for i in range(warm + repeats):
        t0 = time.perf_counter()
        fn(v, o, m, g, *geom())
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    return times, torch.get_num_threads(), kind, desc


def run_reference_arm(args, rank, world):
    """`--impl reference`: the reference's CPU implementation of the path on ALL images of the workload
    (N = 16, same shape: same_config), every host thread, rank 0 only."""
    if rank != 0:
        return
    n = CFG["N"]
    times, threads, kind, desc = cpu_reference_run(n, args.steps, warm=args.warmup)
    total = sum(times)
    val = points_per_step(n) * len(times) / total
    sample = f"all {n} images per step (the full workload, fp32 arithmetic on bf16-rounded inputs), {desc}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample, "same_config": True},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------- next row (8f.1)
def proj_row(dev, dtype):
    """Fused offset/mask projection vs eager at the bench shape; CUDA-graph replays so that the
    numbers are kernel time, inputs rotating over 4 buffers (> L2)."""
    import torch.nn.functional as F
    from yolo_somi_b200.ops_dcnv3.functions import offset_mask_proj as omp
    M, C, G = CFG["N"] * CFG["H"] * CFG["W"], CFG["C"], CFG["G"]
    if not omp.eligible(torch.empty(1, C, device=dev), G, CFG["K"] ** 2, dtype):
        return None
    g = torch.Generator(device="cpu").manual_seed(7)
    xs = [torch.randn(M, C, generator=g).to(dtype).to(dev) for _ in range(4)]
    w_off = (torch.randn(2 * G * 9, C, generator=g) / 16).to(dtype).to(dev)
    w_msk = (torch.randn(G * 9, C, generator=g) / 8).to(dtype).to(dev)
    b_off, b_msk = torch.randn(2 * G * 9, generator=g).to(dtype).to(dev), torch.randn(G * 9, generator=g).to(dtype).to(dev)

    def fused(x):
        with torch.no_grad():
            return omp.OffsetMaskProj.apply(x, w_off, b_off, w_msk, b_msk, G, dtype)

    def eager(x):   # models/ops_dcnv3/modules/dcnv3.py:330-334
        with torch.no_grad():
            o = F.linear(x, w_off, b_off)
            m = F.softmax(F.linear(x, w_msk, b_msk).reshape(M, G, -1).float(), -1).reshape(M, -1).to(dtype)
            return o, m

    def graph_us(fn, reps=8, replays=6):
        for i in range(3):
            fn(xs[i % 4])
        torch.cuda.synchronize()
        gr, st = torch.cuda.CUDAGraph(), torch.cuda.Stream()
        with torch.cuda.stream(st):
            with torch.cuda.graph(gr):
                for i in range(reps):
                    fn(xs[i % 4])
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(replays):
            gr.replay()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / (reps * replays)

    of, mf = fused(xs[0]); oe, me = eager(xs[0])
    err = max(float((of.float() - oe.float()).abs().max()), float((mf.float() - me.float()).abs().max()))
    t_f, t_e = graph_us(fused), graph_us(eager)
    by = 2 * (M * C + M * 3 * G * 9) + 2 * 3 * G * 9 * C
    peak, _ = peaks()
    return {"fused_us": t_f, "eager_us": t_e, "speedup": t_e / t_f, "algorithmic_bytes": by,
            "achieved_gbs": by / t_f / 1e3, "hbm_frac": by / t_f / 1e3 / peak,
            "tflops": 2.0 * M * C * 3 * G * 9 / t_f / 1e6, "max_abs_diff_vs_eager": err,
            "what": "offset + mask linears + softmax over the 9 points (modules/dcnv3.py:330-334) as one "
                    "tcgen05 GEMM with TMEM accumulators and a bias/softmax/cast epilogue"}


def layer_row(dev, dtype):
    """The whole DCNv3 layer (modules/dcnv3.py:222-379) forward + backward at the bench shape: the
    mirrored layer with both fused producers (dwconv+LN+GELU, offset/mask projection) on, and with
    them off (the reference's own sequence of PyTorch ops around the same sampler; the bias
    gradients of the tall projections are ones-row GEMMs in both columns)."""
    from yolo_somi_b200.ops_dcnv3.modules import DCNv3 as Layer
    torch.manual_seed(0)
    layer = Layer(channels=CFG["C"], group=CFG["G"], kernel_size=CFG["K"], offset_scale=CFG["sigma"]).to(dev).to(dtype)
    with torch.no_grad():
        layer.offset.weight.normal_(0, 0.02); layer.mask.weight.normal_(0, 0.1)
    xs = [torch.randn(CFG["N"], CFG["H"], CFG["W"], CFG["C"], device=dev, dtype=dtype, requires_grad=True) for _ in range(2)]
    go = torch.randn(CFG["N"], CFG["H"], CFG["W"], CFG["C"], device=dev, dtype=dtype)

    def timed(n=30, fn=None):
        fn = fn or layer
        for i in range(6):
            fn(xs[i % 2]).backward(go)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(xs[i % 2]).backward(go)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    fused_eager_ms = timed()
    old = {k: os.environ.get(k) for k in ("DCNV3_FUSED_PROJ", "DCNV3_FUSED_DWCONV")}
    os.environ["DCNV3_FUSED_PROJ"] = "0"; os.environ["DCNV3_FUSED_DWCONV"] = "0"
    try:
        unfused_ms = timed()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    # eager, the ~75 launches of the layer's forward + backward take the host as long to enqueue as the GPU to run
    # (scripts/layer_row_probe.py: 1.08 - 1.21 ms of enqueue against 1.15 - 1.26 ms): the row is the GPU's time,
    # measured on the captured form (train_step.graph_block = make_graphed_callables, which patches the module's
    # forward in place -- hence last), with the eager time beside it
    fused_ms, graphed = fused_eager_ms, False
    try:
        from yolo_somi_b200.train_step import graph_block
        gl = graph_block(layer, (xs[0].detach().clone().requires_grad_(True),))
        fused_ms, graphed = min(fused_eager_ms, timed(fn=gl)), True
        del gl
    except Exception:
        pass
    return {"fwd_bwd_ms_fused_producers": fused_ms, "fwd_bwd_ms_fused_producers_eager": fused_eager_ms,
            "fused_row_from_cuda_graph": graphed, "fwd_bwd_ms_pytorch_producers": unfused_ms,
            "what": "DCNv3 layer forward + backward (input_proj, dwconv+LN+GELU, offset/mask, sampler, "
                    "output_proj), N=16 80x80 C=256 G=16 bf16; sampler = this library in both columns"}


# ------------------------------------------------------------------------------- our arm
# ------------------------------------------------------------------------------- BASELINE configs[2] / [3]
def cfg5_row(dev, iters=10):
    """BASELINE configs[4]: the high-resolution sweep -- DCNv3 at 192 x 192 (1536 x 1536 input, stride 8), C = 256,
    group 8 / 16 / 32, bf16, N = 4 images, forward + backward throughput and the HBM-roofline fraction of
    the algorithmic bytes s (3V + 2O + 9Q).  Two rotating input sets per group count (each > the 126 MB L2)."""
    import DCNv3
    n, h, w, c = 4, 192, 192, 256
    peak, _ = peaks()
    rows = {}
    for G in (8, 16, 32):
        geo = (3, 3, 1, 1, 1, 1, 1, 1, G, c // G, 1.0)
        sets = []
        for i in range(2):
            g = torch.Generator(device="cpu").manual_seed(50 + G + i)
            value = torch.randn(n, h, w, c, generator=g)
            offset = torch.randn(n, h, w, G * 18, generator=g)
            mask = torch.softmax(torch.randn(n, h, w, G, 9, generator=g), -1).reshape(n, h, w, G * 9)
            grad = torch.randn(n, h, w, c, generator=g)
            sets.append(tuple(t.to(torch.bfloat16).to(dev) for t in (value, offset, mask, grad)))
        for v, o, m, go in sets:
            DCNv3.dcnv3_forward(v, o, m, *geo, 256); DCNv3.dcnv3_backward(v, o, m, *geo, go, 256)
        a, b_, c_ = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        a.record()
        for i in range(iters):
            v, o, m, go = sets[i % 2]
            DCNv3.dcnv3_forward(v, o, m, *geo, 256)
        b_.record()
        for i in range(iters):
            v, o, m, go = sets[i % 2]
            DCNv3.dcnv3_backward(v, o, m, *geo, go, 256)
        c_.record()
        torch.cuda.synchronize()
        f, bw = a.elapsed_time(b_) / iters, b_.elapsed_time(c_) / iters
        pts = n * h * w * G * 9
        V = n * h * w * c
        nbytes = 2 * (3 * V + 2 * V + 9 * pts)
        rows[f"G{G}"] = {"fwd_ms": f, "bwd_ms": bw, "points": pts, "pts_per_s": pts / ((f + bw) * 1e-3),
                         "algorithmic_bytes": nbytes, "hbm_frac": nbytes / ((f + bw) * 1e-3) / 1e9 / peak}
        del sets
        torch.cuda.empty_cache()
    return {"what": "BASELINE configs[4]: DCNv3 core fwd+bwd at 192x192 (1536x1536 input), C=256, group 8/16/32, bf16, N=4, "
                    "CUDA events, %d calls per pass over two rotating input sets" % iters, "rows": rows}


def _kernel_split(prof):
    """CUDA time of one profiled window by kernel family (torch.profiler / kineto, rank 0)."""
    fam = {"nccl": 0.0, "dcnv3_sm100": 0.0, "conv_gemm": 0.0, "batchnorm_layernorm": 0.0, "elementwise_act_loss": 0.0,
           "copy_cat_permute": 0.0, "optimizer_ema": 0.0, "other": 0.0}
    names = {"nccl": set(), "dcnv3_sm100": set(), "other": set()}
    ours = ("dcnv3", "fwd_gs", "bwd_dots", "bwd_vmma", "bwd_strip", "bwd_scatter", "bwd_mma", "fwd_tile", "fwd_gather",
            "narrow_f32", "offset_mask_proj", "dwconv_ln_gelu", "absmax", "narrow_fixed")
    for ev in prof.events():
        if getattr(ev, "device_type", None) is None or "cuda" not in str(ev.device_type).lower():
            continue
        t = float(getattr(ev, "device_time", 0.0) or getattr(ev, "cuda_time", 0.0) or 0.0)
        n = ev.name
        low = n.lower()
        if "nccl" in low:
            fam["nccl"] += t; names["nccl"].add(n.split("(")[0][:60])
        elif any(k in n for k in ours):
            fam["dcnv3_sm100"] += t; names["dcnv3_sm100"].add(n.split("<")[0].split("(")[0][-40:])
        elif any(k in low for k in ("gemm", "conv", "cudnn", "cutlass", "xmma", "sm90", "sm100", "nvjet", "wgrad", "dgrad")):
            fam["conv_gemm"] += t
        elif any(k in low for k in ("batch_norm", "batchnorm", "layer_norm", "layernorm", "bn_fw", "bn_bw")):
            fam["batchnorm_layernorm"] += t
        elif any(k in low for k in ("multi_tensor", "foreach", "lpnorm")):
            fam["optimizer_ema"] += t
        elif any(k in low for k in ("copy", "cat", "permute", "transpose", "memcpy", "memset", "index", "scatter", "gather", "upsample")):
            fam["copy_cat_permute"] += t
        elif any(k in low for k in ("elementwise", "silu", "sigmoid", "gelu", "reduce", "softmax", "binary_cross", "pool")):
            fam["elementwise_act_loss"] += t
        else:
            fam["other"] += t; names["other"].add(n.split("<")[0].split("(")[0][-50:])
    return fam, {k: sorted(v)[:8] for k, v in names.items()}


def train_row(args, rank, world, local_rank, dev, amp_dtype=torch.float16, graph=False, weak=False):
    """BASELINE configs[3]: the YOLOv5l-DCNv3 training step on synthetic VisDrone-shaped 640 x 640 batches, GLOBAL batch
    128 split over the ranks (strong scaling: 128 / 64 / 32 / 16 images per GPU at 1 / 2 / 4 / 8), bf16 autocast,
    DistributedDataParallel over NCCL with one gradient all-reduce per optimizer step, SGD + fused EMA.  Every rank
    takes part; returns the row on rank 0.  configs[2] (inference, batch 32, one GPU) is measured at N = 1."""
    from yolo_somi_b200.train_step import FusedModelEMA, TrainStep, make_optimizer, synthetic_batch, wrap_ddp
    from yolo_somi_b200.yolov5l_dcnv3 import YOLOv5lDCNv3, layers
    gb = int(os.environ.get("BENCH_TRAIN_BATCH", 128))
    if gb % world:
        return None
    per = gb // world
    if weak:                       # the same per-GPU batch at every N: global batch 128 x N
        per, gb = gb, gb * world
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True
    model = YOLOv5lDCNv3(nc=10).to(dev).to(memory_format=torch.channels_last)
    n_params = sum(p.numel() for p in model.parameters())
    ema = FusedModelEMA(model) if rank == 0 else None
    side = torch.cuda.Stream(dev) if graph else None
    ddp = wrap_ddp(model, local_rank, stream=side)
    ts = TrainStep(ddp, nc=10, optimizer=make_optimizer(model), ema=ema, autocast_dtype=amp_dtype, graph=graph, graph_stream=side)
    # the reference's loader hands uint8 images on the host; the step copies them (train.py:249-250)
    imgs, targets = synthetic_batch(per, 640, device="cpu", seed=rank)
    host_u8 = (imgs * 255).to(torch.uint8).pin_memory()
    host_t = targets.pin_memory()

    def one():
        x = host_u8.to(dev, non_blocking=True).to(memory_format=torch.channels_last).float().div_(255)
        t = host_t.to(dev, non_blocking=True)
        return ts.step(x, t)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    steps = int(os.environ.get("BENCH_TRAIN_STEPS", 6))
    # graph mode: `graph_after` eager steps (11 under DDP), the capturing step, one replay -- all before the timed region
    n_warm = ts.graph_after + 2 if graph else 3
    for _ in range(n_warm):
        one()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = one()
    e1.record()
    barrier()
    from yolo_somi_b200.sharding import max_over_ranks
    ms = max_over_ranks(e0.elapsed_time(e1), dev) / steps
    loss_val = float(loss)
    # where the step's GPU time goes (rank 0): NCCL all-reduce vs this library's kernels vs cuDNN / cuBLAS
    split = kernels = None
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(2):
                one()
            torch.cuda.synchronize()
        if rank == 0:
            fam, kernels = _kernel_split(prof)
            tot = sum(fam.values()) or 1.0
            split = {k: {"ms_per_step": v / 2e3, "share": v / tot} for k, v in fam.items()}
    except Exception as exc:  # profiling is evidence, not the measurement
        split = {"error": str(exc)[:200]}
    barrier()
    infer = None
    if world == 1 and not graph and not weak:
        m = ema.ema
        x = torch.rand(32, 3, 640, 640, device=dev).to(memory_format=torch.channels_last)
        with torch.no_grad(), torch.autocast("cuda", dtype=amp_dtype):
            for _ in range(3):
                m(x)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                m(x)
            b.record(); torch.cuda.synchronize()
        infer = {"img_per_s": 32 * 10 / (a.elapsed_time(b) * 1e-3), "batch": 32, "ms_per_batch": a.elapsed_time(b) / 10,
                 "config": f"BASELINE configs[2]: YOLOv5l-DCNv3 inference, synthetic 640x640 batch 32, {str(amp_dtype)[6:]} autocast, EMA weights"}
        # the deployment form: BatchNorms folded (the reference's fuse() before val / detect) and the forward replayed
        # as one CUDA graph (hosting.fuse_for_inference, hosting.GraphedInference); the eager number stays beside it
        try:
            import copy
            from yolo_somi_b200.hosting import GraphedInference, fuse_for_inference
            fm = fuse_for_inference(copy.deepcopy(m), half=(amp_dtype == torch.float16))
            xg = x.half() if amp_dtype == torch.float16 else x
            gi = GraphedInference(fm, xg, autocast_dtype=None if amp_dtype == torch.float16 else amp_dtype)
            for _ in range(3):
                gi(xg)
            torch.cuda.synchronize()
            a.record()
            for _ in range(10):
                gi(xg)
            b.record(); torch.cuda.synchronize()
            infer["eager"] = {"img_per_s": infer["img_per_s"], "ms_per_batch": infer["ms_per_batch"]}
            g_ips = 32 * 10 / (a.elapsed_time(b) * 1e-3)
            infer["fused_bn_cuda_graph"] = {"img_per_s": g_ips, "ms_per_batch": a.elapsed_time(b) / 10}
            if g_ips > infer["img_per_s"]:
                infer["img_per_s"], infer["ms_per_batch"] = g_ips, a.elapsed_time(b) / 10
                infer["mode"] = ("BatchNorms folded (conv weights / output_proj; the folded bias joins the activation in one pass of "
                                 "dcnv3_bias_act_sm100), model.half() as the reference's val.py, forward replayed as one CUDA graph")
            del gi, fm
        except Exception as exc:
            infer["fused_bn_cuda_graph"] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
    mem = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    graph_state = {"requested": bool(graph), "captured": ts._g is not None, "error": ts.graph_error, "eager_steps_before_capture": ts.graph_after}
    del ts, ddp, model, ema
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    grad_bytes = n_params * 4
    limiter = None
    if split and "error" not in split:
        limiter = max(split, key=lambda k: split[k]["share"])
    return {
        "img_per_s": gb / (ms * 1e-3), "ms_per_step": ms, "global_batch": gb, "per_gpu_batch": per, "n_gpus": world,
        "scaling": "weak" if weak else "strong", "steps": steps, "warmup": n_warm, "loss_last": loss_val, "peak_mem_gib": mem,
        "cuda_graph": graph_state,
        "model": "YOLOv5l-DCNv3 (stock YOLOv5l layout, the four head C3 stages are C3_DCNv3: 12 DCNv3 layers)",
        "params": n_params, "layers": layers(),
        "amp_dtype": str(amp_dtype)[6:],
        "step": f"uint8 host batch -> H2D -> {str(amp_dtype)[6:]} autocast forward -> YOLO-shaped surrogate loss -> "
                "(GradScaler for fp16, as train.py:263-274) backward -> unscale -> clip -> fused SGD(nesterov), which "
                "skips an overflowed step on the device -> fused EMA (rank 0); no host sync inside the step",
        "ddp": {"backend": "nccl" if world > 1 else None, "gradient_as_bucket_view": True, "static_graph": True,
                "broadcast_buffers": False, "bucket_cap_mb": 64, "allreduce_bytes_per_step": grad_bytes if world > 1 else 0},
        "gpu_time_split_rank0": split, "kernels_seen": kernels, "limiter": limiter, "inference": infer,
        "config": "BASELINE configs[3]: YOLO-SOMI training step, synthetic VisDrone-shaped 640x640, global batch %d" % gb,
    }


def run_ours(args, rank, world, local_rank):
    import DCNv3  # the drop-in shim (repo root) over libdcnv3_sm100.so; raises if the .so is absent
    from yolo_somi_b200 import _native
    _native.load()

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dtype = torch.bfloat16
    n = CFG["N"]
    g = geom()
    sets = [make_inputs(n, dev, dtype, seed=100 * rank + i) for i in range(ROTATE)]

    def step(i):
        v, o, m, go = sets[i % ROTATE]
        out = DCNv3.dcnv3_forward(v, o, m, *g, 256)
        grads = DCNv3.dcnv3_backward(v, o, m, *g, go, 256)
        return out, grads

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        # per-pass timing (same launches, same rotation) for the roofline object: events are
        # recorded back to back and read after one synchronize, so no host latency is included
        marks = []
        for i in range(args.steps):
            v, o, m, go = sets[i % ROTATE]
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record(); DCNv3.dcnv3_forward(v, o, m, *g, 256)
            b.record(); DCNv3.dcnv3_backward(v, o, m, *g, go, 256)
            c.record()
            marks.append((a, b, c))
        torch.cuda.synchronize()
        fwd_ms = [a.elapsed_time(b) for a, b, c in marks]
        bwd_ms = [b.elapsed_time(c) for a, b, c in marks]
    clocks = clk.summary()
    from yolo_somi_b200.sharding import max_over_ranks
    ms = max_over_ranks(ms, dev)

    # ---- end to end: pinned host buffers in, results back to pinned host buffers, through the
    # host-buffer entry point of the C ABI (dcnv3_host_pipeline_*): every step copies its four
    # inputs H2D and its four results D2H; chunks of 8 images overlap copy-in, kernels and copy-out
    from yolo_somi_b200.host_pipeline import DCNv3HostPipeline
    cpu_bind = None
    if world > 1 and hasattr(os, "sched_setaffinity"):
        # one slice of the host cores per rank BEFORE the pinned buffers are allocated and first touched (the copies'
        # staging threads and the pages then stay with the rank; every GPU of this pool reports NUMA node 0, so this
        # cannot fix a shared root complex -- the bare-copy baseline below says what the box can do)
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            mine = cores[local_rank * per:(local_rank + 1) * per] or cores
            os.sched_setaffinity(0, mine)
            cpu_bind = [mine[0], mine[-1]]
        except OSError:
            cpu_bind = None
    host_in = [t.cpu().pin_memory() for t in sets[0]]
    pipe = DCNv3HostPipeline(CFG["H"], CFG["W"], CFG["G"], CFG["C"] // CFG["G"], kernel=CFG["K"],
                             stride=CFG["stride"], pad=CFG["pad"], dilation=CFG["dil"],
                             offset_scale=CFG["sigma"], dtype=dtype, chunk_images=int(os.environ.get("BENCH_E2E_CHUNK", 8)), device=dev)
    sv, so, sm, sy = pipe.shapes(n)
    host_out = [torch.empty(shp, dtype=dtype).pin_memory() for shp in (sy, sv, so, sm)]
    def e2e_step():
        pipe.run(*host_in, *host_out)
    # as many steps as the device-resident leg (capped at 50): over ten steps the pipeline's fill and drain -- the first
    # chunk's H2D and the last chunk's D2H overlap nothing -- were a tenth of the measured time
    e2e_steps = max(3, min(args.steps, 50))
    for _ in range(2):
        e2e_step()
    pipe.sync()
    barrier()
    t0 = time.perf_counter()            # the pipeline runs on its own streams: host clock around
    for _ in range(e2e_steps):          # enqueue + full drain (sync waits for the last D2H)
        e2e_step()
    pipe.sync()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    e2e_ms = max_over_ranks(e2e_ms, dev)
    h2d = sum(t.numel() * t.element_size() for t in host_in)
    d2h = sum(t.numel() * t.element_size() for t in host_out)
    # the pipelined results must be the device-resident path's results
    ref_out = DCNv3.dcnv3_forward(*sets[0][:3], *g, 256)
    if not torch.equal(ref_out.cpu(), host_out[0]):
        raise SystemExit("bench.py: host pipeline output differs from the device-resident forward")
    pipe.close()
    # the box's ceiling for this traffic: the same bytes as bare cudaMemcpyAsync calls (one per tensor), H2D on one
    # stream and D2H on another at the same time, no kernels -- every rank at once, max over ranks
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    dev_in = [torch.empty_like(t, device=dev) for t in host_in]
    dev_out = [torch.empty_like(t, device=dev) for t in host_out]
    def bare_copy():
        with torch.cuda.stream(s_in):
            for h, d in zip(host_in, dev_in):
                d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s_out):
            for h, d in zip(host_out, dev_out):
                h.copy_(d, non_blocking=True)
    bare_copy(); torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        bare_copy()
    torch.cuda.synchronize()
    bare_ms = (time.perf_counter() - t0) * 1e3
    barrier()
    bare_ms = max_over_ranks(bare_ms, dev)
    del dev_in, dev_out

    # ---- SURVEY 8(f) rank 1, measured beside the sampler: the fused offset/mask projection
    # (tcgen05 GEMM + bias + softmax epilogue) against the layer's two linears + softmax, same shape
    def timed_mode(env, iters=5):
        for k, val in env.items():
            os.environ[k] = val
        try:
            v, o, m, go = sets[0]
            for _ in range(2):
                DCNv3.dcnv3_forward(v, o, m, *g, 256); DCNv3.dcnv3_backward(v, o, m, *g, go, 256)
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            for i in range(iters):
                DCNv3.dcnv3_forward(*sets[i % ROTATE][:3], *g, 256)
            b.record()
            for i in range(iters):
                vv, oo, mm, gg = sets[i % ROTATE]
                DCNv3.dcnv3_backward(vv, oo, mm, *g, gg, 256)
            c.record(); torch.cuda.synchronize()
            return {"fwd_ms": a.elapsed_time(b) / iters, "bwd_ms": b.elapsed_time(c) / iters}
        finally:
            for k in env:
                os.environ.pop(k, None)
    modes = None
    if rank == 0:
        modes = {
            "default": "forward: bilinear x mask coefficients rounded to bf16 (one FHFMA per element); backward: exact "
                       "grad_out x value products, grad_value through the tcgen05 product with bf16 coefficient sums; "
                       "fp32 accumulation everywhere (errors: profiles/parity_r2.json)",
            "strict_fp32_coefficients": dict(timed_mode({"DCNV3_WEIGHTS": "split", "DCNV3_BWD": "scatter"}),
                                             env="DCNV3_WEIGHTS=split DCNV3_BWD=scatter",
                                             note="every coefficient fp32 as the reference's opmath_t; errors = output rounding only"),
            "deterministic": dict(timed_mode({"DCNV3_DETERMINISTIC": "1"}), env="DCNV3_DETERMINISTIC=1",
                                  note="bit-reproducible backward: 64-bit fixed-point accumulation of grad_value"),
        }
    proj = proj_row(dev, dtype) if rank == 0 else None
    layer = layer_row(dev, dtype) if rank == 0 else None
    # ---- the reference's own CUDA kernels recompiled for sm_100a (baseline/_ref/DCNv3_refcuda.so, built by
    # scripts/build_reference_cuda.py), timed beside this library on this box: fp16 / fp32, it has no bf16 dispatch
    cfg5 = None
    if rank == 0:
        try:
            cfg5 = cfg5_row(dev)
        except Exception as exc:
            cfg5 = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
    refcuda = None
    if rank == 0 and world == 1 and os.environ.get("BENCH_REF_CUDA", "1") != "0":
        try:
            sys.path.insert(0, str(ROOT / "scripts"))
            import bench_reference_cuda
            refcuda = bench_reference_cuda.run(iters=5, dev=dev)
        except Exception as exc:
            refcuda = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
    # ---- BASELINE configs[2] / [3]: the model-level step, every rank takes part (NCCL gradient all-reduce)
    del sets, host_in, host_out
    torch.cuda.empty_cache()
    train_on = not args.no_train and int(os.environ.get("BENCH_TRAIN_BATCH", 128)) % world == 0
    train = train_row(args, rank, world, local_rank, dev) if train_on else None
    if train is not None and world == 1 and os.environ.get("BENCH_TRAIN_BF16", "1") != "0":
        # the same step in bf16: PyTorch 2.11 has no cuDNN BatchNorm for bf16 (native kernel: 2.9x slower,
        # scripts/bn_probe.py), which is why the headline row uses the reference's own AMP dtype
        alt = train_row(args, rank, world, local_rank, dev, amp_dtype=torch.bfloat16)
        train["bf16_autocast"] = {k: alt[k] for k in ("img_per_s", "ms_per_step", "limiter", "gpu_time_split_rank0")}

    # ---- the same step as ONE captured CUDA graph (TrainStep(graph=True)), and at N > 1 the weak-scaling row (128 images
    # per GPU).  Both run last and under a watchdog: a capture that hangs (NCCL inside a capture) must not cost the line.
    extra_on = train_on and os.environ.get("BENCH_TRAIN_GRAPH", "1") != "0"

    def finish(graph_row, weak_row):
        if graph_row is not None and train is not None:
            keep = ("img_per_s", "ms_per_step", "cuda_graph", "peak_mem_gib", "loss_last", "gpu_time_split_rank0", "limiter")
            train["whole_step_cuda_graph"] = {k: graph_row.get(k) for k in keep} if "error" not in graph_row else graph_row
            train["eager"] = {"img_per_s": train["img_per_s"], "ms_per_step": train["ms_per_step"]}
            st = graph_row.get("cuda_graph") or {}
            if st.get("captured") and not st.get("error") and graph_row["img_per_s"] > train["img_per_s"]:
                train["img_per_s"], train["ms_per_step"] = graph_row["img_per_s"], graph_row["ms_per_step"]
                train["mode"] = "whole step replayed as one CUDA graph (TrainStep(graph=True)); the eager step beside it"
            else:
                train["mode"] = "eager"
        if weak_row is not None and train is not None:
            train["weak_scaling_128_per_gpu"] = {k: weak_row.get(k) for k in ("img_per_s", "ms_per_step", "global_batch", "per_gpu_batch", "limiter", "cuda_graph")} \
                if "error" not in weak_row else weak_row
        pts = points_per_step()
        value = world * pts * args.steps / (ms * 1e-3)
        ab = algorithmic_bytes()
        peak, peak_src = peaks()
        traffic, traffic_src = measured_traffic()
        bwd_t = statistics.mean(bwd_ms) * 1e-3
        fwd_t = statistics.mean(fwd_ms) * 1e-3
        ach = ab["bwd"] / bwd_t / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": n, "points_per_step_per_gpu": pts,
                       "l2": f"{ROTATE} rotating input sets (each step's operands ~390 MB > 126 MB L2)",
                       "arithmetic": "fp32 accumulate, bf16 I/O, sampling coefficients rounded to bf16 (defaults)"},
            "clocks": clocks,
            "e2e": {"value": world * pts * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / e2e_steps,
                    "api": "dcnv3_host_pipeline_run (C ABI, pinned host buffers, 8-image chunks, H2D | kernels | D2H on three streams)",
                    "timing": "host clock around enqueue + drain of all steps (the pipeline owns its streams)",
                    "bare_copy_ms_per_step": bare_ms / e2e_steps,
                    "bare_copy": "the same H2D + D2H bytes as plain cudaMemcpyAsync per tensor on two streams, no kernels, all ranks at once",
                    "pcie_gbs_each_way_per_gpu_bare": h2d / (bare_ms / e2e_steps * 1e-3) / 1e9,
                    "host_link_gbs_aggregate_bare": world * (h2d + d2h) / (bare_ms / e2e_steps * 1e-3) / 1e9,
                    "pipeline_over_bare_copy": (e2e_ms / e2e_steps) / (bare_ms / e2e_steps),
                    "cpu_cores_bound": cpu_bind},
            "gpu_launches": 6 * args.steps,   # per step: fwd_gs | bwd_dots, bwd_vres, far_points + the two conditional fall-back launches (bwd_vmma, narrow_f32: they exit at once)
            "roofline": {"bound": "hbm", "kernel": "backward pass: bdots::bwd_dots (channel sums) + vres::bwd_vres (grad_value, tcgen05 with the accumulator resident in TMEM, written once) + vres::far_points (+ two conditional fall-back launches that exit at once), chained by programmatic dependent launch",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes": ab["bwd"], "avg_ms": bwd_t * 1e3},
            "passes": {"fwd_ms": fwd_t * 1e3, "bwd_ms": bwd_t * 1e3,
                       "fwd_gbs": ab["fwd"] / fwd_t / 1e9, "fwd_frac": ab["fwd"] / fwd_t / 1e9 / peak,
                       "step_gbs": (ab["fwd"] + ab["bwd"]) / (ms / args.steps * 1e-3) / 1e9,
                       "step_frac": (ab["fwd"] + ab["bwd"]) / (ms / args.steps * 1e-3) / 1e9 / peak},
        }
        if proj is not None:
            line["next_rows"] = {"offset_mask_proj": proj, "layer": layer}
        if modes is not None:
            line["precision_modes"] = modes
        if refcuda is not None:
            line["reference_cuda_kernels"] = refcuda
        if cfg5 is not None:
            line["cfg5_sweep"] = cfg5
        if train is not None:
            line["train_step"] = train
            line["img_per_s"] = train["img_per_s"]
        if world == 1 and not args.no_cpu:
            # the CPU reference beside the GPU number, same run, same box: the full workload (all 16 images), 5 passes
            times, threads, kind, desc = cpu_reference_run(CFG["N"], 5)
            line["cpu_baseline"] = {
                "value": points_per_step() / min(times), "unit": UNIT, "cores": threads, "kind": kind,
                "sample": f"all {CFG['N']} images (the full workload), fp32 on bf16-rounded inputs, best of 5 after 1 warm-up "
                          f"({sum(times):.1f} s of CPU work), {desc}"}
        print(json.dumps(line))

    graph_row = weak_row = None
    if extra_on:
        import threading
        limit = int(os.environ.get("BENCH_TRAIN_GRAPH_TIMEOUT", 300))
        done = threading.Event()

        def bail():
            if done.is_set():
                return
            if rank == 0:
                msg = {"error": "did not finish within %d s (watchdog); the eager rows stand" % limit}
                finish(msg if graph_row is None else graph_row, msg)
                sys.stdout.flush()
            os._exit(0)
        timer = threading.Timer(limit, bail)
        timer.daemon = True
        timer.start()
        try:
            graph_row = train_row(args, rank, world, local_rank, dev, graph=True)
        except Exception as exc:
            graph_row = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
        if world > 1 and os.environ.get("BENCH_TRAIN_WEAK", "1") != "0":
            try:
                weak_row = train_row(args, rank, world, local_rank, dev, weak=True, graph=True)
            except Exception as exc:
                weak_row = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
        done.set()
        timer.cancel()
    if rank != 0:
        return
    finish(graph_row, weak_row)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-train", action="store_true", help="skip the YOLOv5l-DCNv3 training-step row")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the DCNv3 core has no CPU path); "
                         "use --impl reference for the CPU baseline")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        # NCCL may print its version banner on stdout while the communicator is created; stdout
        # carries exactly one JSON line, so park fd 1 on stderr until the first collective is done
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            torch.distributed.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
