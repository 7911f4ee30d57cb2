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
