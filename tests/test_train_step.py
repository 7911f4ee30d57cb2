"""the training-step utilities around the DCNv3 layers (yolo_somi_b200/train_step.py)."""
import math
import sys
from copy import deepcopy

import pytest
import torch
from torch import nn


class _RefEMA:
    """The reference's ModelEMA.update, restated (utils/torch_utils.py:315-349) -- the checker."""

    def __init__(self, model, decay=0.9999, updates=0):
        self.ema = deepcopy(model).eval()
        self.updates = updates
        self.decay = lambda x: decay * (1 - math.exp(-x / 2000))
        for p in self.ema.parameters():
            p.requires_grad_(False)

    def update(self, model):
        with torch.no_grad():
            self.updates += 1
            d = self.decay(self.updates)
            msd = model.state_dict()
            for k, v in self.ema.state_dict().items():
                if v.dtype.is_floating_point:
                    v *= d
                    v += (1 - d) * msd[k].detach()


def _toy():
    torch.manual_seed(0)
    return nn.Sequential(nn.Conv2d(3, 8, 3), nn.BatchNorm2d(8), nn.SiLU(), nn.Conv2d(8, 4, 1), nn.BatchNorm2d(4))


def test_fused_ema_is_bit_identical_to_the_reference_loop():
    from yolo_somi_b200.train_step import FusedModelEMA
    model = _toy()
    ours, ref = FusedModelEMA(model), _RefEMA(model)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    for step in range(5):
        model.train()
        loss = model(torch.randn(2, 3, 9, 9)).square().mean()     # also moves the BatchNorm buffers
        opt.zero_grad(); loss.backward(); opt.step()
        ours.update(model); ref.update(model)
    assert ours.updates == ref.updates == 5
    a, b = ours.ema.state_dict(), ref.ema.state_dict()
    assert a.keys() == b.keys()
    for k in a:
        assert torch.equal(a[k], b[k]), k                         # floating entries AND the integer counters
    assert not ours.ema.training and not any(p.requires_grad for p in ours.ema.parameters())
    # attributes follow as in the reference's copy_attr
    model.nc, model._private = 10, 1
    ours.update_attr(model, include=["nc"])
    assert ours.ema.nc == 10 and not hasattr(ours.ema, "_private")


def test_fused_ema_takes_a_wrapped_model():
    from yolo_somi_b200.train_step import FusedModelEMA
    model = _toy()
    wrapped = nn.DataParallel(model)
    ema = FusedModelEMA(wrapped)
    assert not isinstance(ema.ema, nn.DataParallel)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(1.0)
    ema.update(wrapped)
    d = 0.9999 * (1 - math.exp(-1 / 2000))
    p0, e0 = next(model.parameters()), next(ema.ema.parameters())
    assert torch.allclose(e0, (p0 - 1.0) * d + (1 - d) * p0)
    # the cached tensor lists follow a model whose storage moved (.double() re-allocates every tensor)
    model.double()
    before = next(ema.ema.parameters()).clone()
    ema.update(wrapped)
    d2 = 0.9999 * (1 - math.exp(-2 / 2000))
    assert torch.allclose(next(ema.ema.parameters()), before * d2 + ((1 - d2) * next(model.parameters())).float())


def test_reference_import_paths_resolve_to_this_library():
    """Checkpoints of the reference pickle modules by class path (train.py:309-323): the paths must import."""
    from yolo_somi_b200.train_step import install_reference_aliases
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "models" or k.startswith("models.")}
    try:
        done = install_reference_aliases(force=True)
        assert "models.ops_dcnv3.modules.dcnv3" in done
        import importlib
        mod = importlib.import_module("models.ops_dcnv3.modules.dcnv3")
        from yolo_somi_b200.ops_dcnv3.modules import dcnv3 as ours
        assert mod is ours and hasattr(mod, "DCNv3")
        fn = importlib.import_module("models.ops_dcnv3.functions.dcnv3_func")
        assert hasattr(fn, "DCNv3Function")
        # every public name of the reference's two packages (functions/__init__.py, modules/__init__.py) resolves,
        # and the debug layer's class path carries the same state_dict keys as the layer
        from models.ops_dcnv3.functions import DCNv3Function, dcnv3_core_pytorch  # noqa: F401
        from models.ops_dcnv3.modules import DCNv3, DCNv3_pytorch
        a, b = DCNv3(channels=32, group=2), DCNv3_pytorch(channels=32, group=2)
        assert list(a.state_dict()) == list(b.state_dict()) and isinstance(b, DCNv3)
    finally:
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.modules.update(saved)


@pytest.mark.gpu
def test_dcnv3_layer_runs_under_make_graphed_callables():
    """A static-shape DCNv3 layer (fused producers, split backward with its side stream) captured by
    torch.cuda.make_graphed_callables gives the eager forward and gradients."""
    from yolo_somi_b200.ops_dcnv3.modules import DCNv3
    from yolo_somi_b200.train_step import graph_block
    torch.manual_seed(0)
    layer = DCNv3(channels=128, group=8).cuda().to(torch.bfloat16)
    with torch.no_grad():                                   # offsets / masks away from their zero init
        layer.offset.weight.normal_(0, 0.02); layer.mask.weight.normal_(0, 0.02)
    # graph the fresh layer first (autograd state created on the default stream beforehand -- e.g. AccumulateGrad
    # nodes of the parameters -- would make the capture depend on the legacy stream); the eager check uses a copy
    eager = deepcopy(layer)
    x = torch.randn(2, 24, 24, 128, device="cuda", dtype=torch.bfloat16)
    graphed = graph_block(layer, (x.clone().requires_grad_(True),))
    x2 = x.clone().requires_grad_(True)
    y2 = graphed(x2)
    gy = torch.randn_like(y2)
    got = torch.autograd.grad(y2, [x2] + list(layer.parameters()), gy)
    x1 = x.clone().requires_grad_(True)
    y = eager(x1)
    want = torch.autograd.grad(y, [x1] + list(eager.parameters()), gy)
    torch.cuda.synchronize()
    assert torch.allclose(y2.float(), y.float(), rtol=2e-2, atol=2e-2 * float(y.float().abs().max()))
    for a, b in zip(got, want):
        scale = float(b.float().abs().max()) + 1e-6
        assert float((a.float() - b.float()).abs().max()) <= 5e-2 * scale


# ----------------------------------------------------------------------------- DDP wiring (gloo, world size 2, CPU)
class _TinyDet(nn.Module):
    """A CPU-runnable stand-in with the step harness's output contract: list of [B, na, H, W, nc + 5] maps."""

    def __init__(self, nc=3):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(3, 8, 3, 2, 1), nn.BatchNorm2d(8), nn.SiLU(), nn.Conv2d(8, 8, 3, 2, 1))
        self.head = nn.Conv2d(8, 2 * (nc + 5), 1)
        self.nc = nc

    def forward(self, x):
        y = self.head(self.body(x))
        b, _, h, w = y.shape
        return [y.view(b, 2, self.nc + 5, h, w).permute(0, 1, 3, 4, 2)]


def _ddp_worker(rank, world, port, out):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from yolo_somi_b200.train_step import FusedModelEMA, TrainStep, make_optimizer, setup_process_group, \
        synthetic_batch, wrap_ddp
    r, lr, w = setup_process_group("gloo")
    try:
        torch.manual_seed(0)
        model = _TinyDet()
        ddp = wrap_ddp(model, lr)
        assert isinstance(ddp, nn.parallel.DistributedDataParallel)
        ts = TrainStep(ddp, nc=3, optimizer=make_optimizer(model, lr=0.1), ema=FusedModelEMA(model) if r == 0 else None,
                       autocast_dtype=None, log_every=1)
        imgs, targets = synthetic_batch(4, 32, nc=3, boxes_per_image=2, device="cpu", seed=r)   # a different shard per rank
        ts.step(imgs, targets, last_micro=True)                   # (static_graph records the graph on a synchronised step)
        ts.step(imgs, targets, last_micro=False)                  # accumulation micro-step: no all-reduce
        g_local = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).clone()
        ts.step(imgs, targets, last_micro=True)                   # all-reduce + optimizer + EMA
        flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        gathered = [torch.zeros_like(flat) for _ in range(w)]
        dist.all_gather(gathered, flat)
        g_all = [torch.zeros_like(g_local) for _ in range(w)]
        dist.all_gather(g_all, g_local)
        out[r] = (bool(torch.equal(gathered[0], gathered[1])), bool(torch.equal(g_all[0], g_all[1])),
                  ts.loss_for_log()[0], float(ts.loss_for_log()[1]))
    finally:
        dist.destroy_process_group()


def test_ddp_wiring_two_ranks_gloo():
    """train.py:165-208 / 263-277 restated (train_step.wrap_ddp / TrainStep): ranks see different shards; a no_sync()
    micro-step leaves the gradients rank-local, the closing micro-step all-reduces once and the parameters stay
    identical on both ranks; the logged loss is available without a per-step synchronisation."""
    import os
    import torch.multiprocessing as mp
    world, port = 2, 29900 + os.getpid() % 300
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_ddp_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    for r in range(world):
        same_params, same_local_grads, step, loss = res[r]
        assert same_params and not same_local_grads
        assert step == 2 and loss == loss and loss > 0


def test_optimizer_groups_follow_the_reference_split():
    from yolo_somi_b200.train_step import make_optimizer
    m = _TinyDet()
    opt = make_optimizer(m)
    decay, no_decay = (g["params"] for g in opt.param_groups)
    assert opt.param_groups[0]["weight_decay"] > 0 and opt.param_groups[1]["weight_decay"] == 0
    assert len(decay) == 3 and len(no_decay) == 5            # 3 conv weights | BN weight + BN bias + 3 conv biases
    assert opt.defaults["nesterov"] and opt.defaults["momentum"] == 0.937


def test_yolov5l_dcnv3_layer_list_and_construction():
    from yolo_somi_b200.yolov5l_dcnv3 import YOLOv5lDCNv3, layers
    spec = layers()
    assert len(spec) == 25 and spec[24][2] == "Detect" and [l[2] for l in spec].count("C3_DCNv3") == 4
    m = YOLOv5lDCNv3(nc=10)
    dcn = m.dcnv3_layers()
    assert len(dcn) == 12 and sorted({(l.channels, l.group) for l in dcn}) == [(128, 8), (256, 16), (512, 32)]
    assert 35e6 < sum(p.numel() for p in m.parameters()) < 50e6


@pytest.mark.gpu
def test_training_step_runs_on_gpu():
    """One real step of the harness on a small input: finite loss, parameters and EMA move, DCNv3 kernels ran."""
    from yolo_somi_b200.train_step import FusedModelEMA, TrainStep, make_optimizer, synthetic_batch
    from yolo_somi_b200.yolov5l_dcnv3 import YOLOv5lDCNv3
    torch.manual_seed(0)
    model = YOLOv5lDCNv3(nc=10).cuda().to(memory_format=torch.channels_last)
    ema = FusedModelEMA(model)
    ts = TrainStep(model, nc=10, optimizer=make_optimizer(model, lr=0.01), ema=ema, log_every=1)
    before = model.model[13].m[0].cv2.dcn.output_proj.weight.detach().clone()
    imgs, targets = synthetic_batch(2, 256, device="cuda")
    for _ in range(2):
        loss = ts.step(imgs.to(memory_format=torch.channels_last), targets)
    torch.cuda.synchronize()
    assert torch.isfinite(loss)
    assert not torch.equal(before, model.model[13].m[0].cv2.dcn.output_proj.weight.detach())
    step, mean_loss = ts.loss_for_log()
    assert step == 2 and mean_loss == mean_loss
    assert ema.updates == 2


@pytest.mark.gpu
@pytest.mark.parametrize("amp", [torch.bfloat16, torch.float16])
def test_whole_step_cuda_graph_follows_the_eager_step(amp):
    """TrainStep(graph=True): three eager steps, a captured step, replays -- losses, parameters and the EMA follow an
    eager TrainStep fed the same batches (atomics in the samplers' backward: close, not bit-equal)."""
    from yolo_somi_b200.train_step import FusedModelEMA, TrainStep, make_optimizer, synthetic_batch
    from yolo_somi_b200.yolov5l_dcnv3 import YOLOv5lDCNv3
    runs = {}
    batches = [synthetic_batch(2, 256, device="cuda", seed=s) for s in range(7)]
    for mode in ("eager", "graph"):
        torch.manual_seed(0)
        model = YOLOv5lDCNv3(nc=10).cuda().to(memory_format=torch.channels_last)
        ema = FusedModelEMA(model)
        ts = TrainStep(model, nc=10, optimizer=make_optimizer(model, lr=0.01), ema=ema, autocast_dtype=amp, log_every=7,
                       graph=(mode == "graph"))
        losses = []
        for imgs, targets in batches:
            losses.append(float(ts.step(imgs.to(memory_format=torch.channels_last), targets)))
        torch.cuda.synchronize()
        if mode == "graph":
            assert ts.graph_error is None and ts._g is not None, ts.graph_error
        assert ema.updates == 7 and ts.loss_for_log()[0] == 7
        runs[mode] = (losses, [p.detach().float().clone() for p in model.parameters()],
                      [p.detach().float().clone() for p in ema.ema.parameters()], ts.loss_for_log()[1])
    le, pe, ee, me = runs["eager"]
    lg, pg, eg, mg = runs["graph"]
    assert all(l == l for l in lg)
    assert max(abs(a - b) for a, b in zip(le, lg)) < 0.05 * max(abs(a) for a in le), (le, lg)
    assert abs(me - mg) < 0.05 * abs(me)
    for a, b in zip(pe, pg):
        assert (a - b).norm() <= 0.05 * a.norm() + 1e-3
    for a, b in zip(ee, eg):
        assert (a - b).norm() <= 0.05 * a.norm() + 1e-3


@pytest.mark.gpu
def test_fused_graphed_inference_matches_the_eager_model():
    """hosting.fuse_for_inference + GraphedInference (BASELINE configs[2] deployment form): same detections as the
    eager eval-mode model with its BatchNorms, within fp16 rounding."""
    import copy
    from yolo_somi_b200.hosting import GraphedInference, fuse_for_inference
    from yolo_somi_b200.yolov5l_dcnv3 import YOLOv5lDCNv3
    torch.manual_seed(0)
    model = YOLOv5lDCNv3(nc=10).cuda().to(memory_format=torch.channels_last).eval()
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.8, 1.2); m.weight.normal_(1, 0.1); m.bias.normal_(0, 0.1)
        for l in model.dcnv3_layers():
            l.offset.weight.normal_(0, 0.02); l.mask.weight.normal_(0, 0.1)
    x = torch.rand(2, 3, 256, 256, device="cuda").to(memory_format=torch.channels_last)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        want = [t.float() for t in model(x)]
    for half in (False, True):                    # autocast over fp32 weights | model.half() as the reference's val.py
        fm = fuse_for_inference(copy.deepcopy(model), half=half)
        xin = x.half() if half else x
        gi = GraphedInference(fm, xin, autocast_dtype=None if half else torch.float16)
        gi(torch.rand_like(xin))                  # another input in between: the replay must take the new one
        got = [t.float().clone() for t in gi(xin)]
        torch.cuda.synchronize()
        assert not any(isinstance(m, torch.nn.BatchNorm2d) for m in gi.model.modules())
        assert any(type(m).__name__ == "BiasAct" for m in gi.model.modules())
        for a, b in zip(got, want):
            assert a.shape == b.shape
            assert (a - b).norm() <= 2e-2 * b.norm() + 1e-3, (half, float((a - b).norm() / b.norm()))


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 64, 7, 9), (1, 1024, 3, 5), (3, 24, 8, 8), (4, 256, 40, 40)])
def test_bias_act_kernel_matches_the_expression(shape, dt):
    """dcnv3_bias_act_sm100 (csrc/dcnv3_hosting.cu) through hosting.BiasAct: act(x + bias) on channels-last 16-bit
    maps, in place, against the fp32 expression rounded once; a map that is not channels-last takes the fallback."""
    from yolo_somi_b200.hosting import BiasAct
    g = torch.Generator(device="cpu").manual_seed(sum(shape))
    x = (3 * torch.randn(*shape, generator=g)).to(dt)
    bias = torch.randn(shape[1], generator=g)
    for kind in ("silu", "identity"):
        mod = BiasAct(bias, kind).cuda()
        t = x.float() + bias.view(1, -1, 1, 1)
        want = (torch.nn.functional.silu(t) if kind == "silu" else t).to(dt)
        xc = x.cuda().to(memory_format=torch.channels_last)
        with torch.no_grad():
            got = mod(xc)
        torch.cuda.synchronize()
        assert got.data_ptr() == xc.data_ptr()                       # in place
        ulp = 2.0 ** -7 if dt == torch.bfloat16 else 2.0 ** -10
        assert float((got.float().cpu() - want.float()).abs().max()) <= ulp * float(want.float().abs().max()) + 1e-6
        with torch.no_grad():
            fb = mod(x.cuda().contiguous())                          # NCHW-contiguous: the PyTorch expression
        assert float((fb.float().cpu() - want.float()).abs().max()) <= 2 * ulp * float(want.float().abs().max()) + 1e-6


def test_train_step_cpu_graph_flag_and_running_mean():
    """graph=True needs CUDA: on the CPU the step stays eager (and accumulation is allowed); the logged loss is the MEAN
    of the steps' losses; the same global batch gives the same update at any world size because the loss is a mean
    (no `loss * WORLD_SIZE`, see TrainStep._work)."""
    from yolo_somi_b200.train_step import TrainStep, make_optimizer, synthetic_batch
    torch.manual_seed(0)
    model = _TinyDet()
    ts = TrainStep(model, nc=3, optimizer=make_optimizer(model, lr=0.05), autocast_dtype=None, log_every=1, graph=True)
    assert ts.graph is False
    imgs, targets = synthetic_batch(4, 32, nc=3, boxes_per_image=2, device="cpu", seed=1)
    losses = [float(ts.step(imgs, targets)) for _ in range(3)]
    ts.step(imgs, targets, last_micro=False)            # allowed in eager mode
    step, mean = ts.loss_for_log()
    assert step == 3 and abs(mean - sum(losses) / 3) < 1e-5 * max(1.0, abs(mean))
    assert losses[2] < losses[0]                         # it trains
