"""Training-step critical path around the DCNv3 layers.

What the reference does once per optimizer step, and what this module offers instead:

* ``ModelEMA.update`` (utils/torch_utils.py:315-349, called at train.py:275-276) walks the whole
  ``state_dict`` in Python and launches two elementwise kernels per floating tensor (``v *= d``,
  ``v += (1 - d) * m``) -- several hundred tiny launches per step for YOLOv5l, all on the training
  stream behind the optimizer.  :class:`FusedModelEMA` keeps the reference's API (``ema``, ``updates``,
  ``decay``, ``update``, ``update_attr``) and arithmetic (same operation order, bit-identical result)
  but updates all tensors of a (device, dtype) with three multi-tensor launches
  (``torch._foreach_mul_`` / ``_foreach_mul`` / ``_foreach_add_``) over tensor lists that are rebuilt only
  when a tensor moved.
* a static-shape forward + backward of a block that contains DCNv3 layers can be captured into CUDA
  graphs: every kernel of this library launches on the current stream, the backward's internal side
  stream is forked and joined with events (legal under capture), and nothing synchronises.
  :func:`graph_block` is ``torch.cuda.make_graphed_callables`` with the warm-up this library needs
  (tensor maps, function attributes and the side stream are created on first use, outside capture).
* checkpoints of the reference pickle whole modules by class path (train.py:309-323,
  models/experimental.py:97-101).  :func:`install_reference_aliases` registers
  ``models.ops_dcnv3.{functions,modules}`` (and the compiled extension's name ``DCNv3``) in
  ``sys.modules`` so that such checkpoints load onto this library's classes, and models saved with this
  library load in a tree that has the reference's package.

Nothing here is on the sampling path; it exists so that the sampler's speed reaches the step time.
"""
from __future__ import annotations

import importlib
import itertools
import math
import sys
from copy import deepcopy
from typing import Callable, Iterable, Sequence

import torch
from torch import nn


def _de_parallel(model: nn.Module) -> nn.Module:
    wrapped = (nn.parallel.DataParallel, nn.parallel.DistributedDataParallel)
    return model.module if isinstance(model, wrapped) else model


class FusedModelEMA:
    """Drop-in for the reference's ``ModelEMA`` (utils/torch_utils.py:315-349) with multi-tensor updates.

    Same semantics: an ``eval()`` deep copy of the (de-parallelised) model, ``updates`` counter,
    ``decay(x) = decay * (1 - exp(-x / 2000))``, every floating-point entry of the ``state_dict``
    (parameters AND buffers) follows ``v = v * d + (1 - d) * m``; integer buffers are left alone exactly
    as in the reference.  The products are formed in the reference's order, so the result is bit-identical
    to the per-tensor loop.
    """

    def __init__(self, model: nn.Module, decay: float = 0.9999, updates: int = 0):
        self.ema = deepcopy(_de_parallel(model)).eval()
        self.updates = updates
        self.decay_base, self.decay_tau = decay, 2000.0
        self.decay = lambda x: decay * (1 - math.exp(-x / 2000))
        for p in self.ema.parameters():
            p.requires_grad_(False)

    def _bind(self, model: nn.Module, ptrs: list) -> None:
        """Group the floating entries of both state_dicts by (device, dtypes): the multi-tensor kernels take
        homogeneous lists.  state_dict() tensors alias the live storage; `ptrs` fingerprints that storage."""
        msd = model.state_dict()
        groups: dict[tuple, tuple[list, list]] = {}
        for k, v in self.ema.state_dict().items():
            if v.dtype.is_floating_point:
                m = msd[k].detach()
                es, ss = groups.setdefault((v.device, v.dtype, m.dtype), ([], []))
                es.append(v)
                ss.append(m)
        self._groups, self._bound = list(groups.values()), (model, ptrs)

    @staticmethod
    def _fingerprint(*mods: nn.Module) -> list:
        return [t.data_ptr() for mod in mods for t in itertools.chain(mod.parameters(), mod.buffers())]

    def update(self, model: nn.Module, device_count: torch.Tensor | None = None) -> None:
        """One EMA update.  ``device_count`` (a 0-dim float64 CUDA tensor holding the number of updates so far) moves
        the decay schedule onto the device: the counter is incremented and ``d`` computed there (fp64, rounded to the
        tensors' fp32 exactly where the host form rounds), so the update can live inside a captured CUDA graph whose
        replays must not bake in one step's decay.  The host counter ``updates`` is the caller's to advance then."""
        model = _de_parallel(model)
        with torch.no_grad():
            if device_count is None:
                self.updates += 1
                d = self.decay(self.updates)
                rest = 1 - d
            else:
                device_count += 1
                d64 = self.decay_base * (1 - torch.exp(-device_count / self.decay_tau))
                d, rest = d64.float(), (1 - d64).float()
            # the tensor lists are rebuilt only when a tensor of either model moved (.to(), new parameters, ...)
            ptrs = self._fingerprint(model, self.ema)
            bound = getattr(self, "_bound", None)
            if bound is None or bound[0] is not model or bound[1] != ptrs:
                self._bind(model, ptrs)
            for es, ss in self._groups:
                torch._foreach_mul_(es, d)                       # v *= d
                scaled = torch._foreach_mul(ss, rest)            # (1 - d) * m   (in m's dtype, as the reference)
                torch._foreach_add_(es, scaled)                  # v += ...

    def update_attr(self, model: nn.Module, include: Sequence[str] = (),
                    exclude: Sequence[str] = ("process_group", "reducer")) -> None:
        """Copy plain attributes, as ``copy_attr`` of the reference does (utils/torch_utils.py)."""
        for k, v in model.__dict__.items():
            if (len(include) and k not in include) or k.startswith("_") or k in exclude:
                continue
            setattr(self.ema, k, v)


def graph_block(block: nn.Module | Callable, sample_args: Iterable[torch.Tensor], warmup: int = 3):
    """CUDA-graph a static-shape block (forward and backward) that contains DCNv3 layers.

    ``torch.cuda.make_graphed_callables`` after ``warmup`` eager forward + backward passes on a side stream, so
    that everything this library creates lazily (tensor maps' driver entry point, function attributes, the
    backward's side stream and events, cuBLAS workspaces of the projections) exists before capture starts.
    Returns the graphed callable; use it in place of ``block`` with tensors of the same shapes and dtypes.
    """
    args = tuple(sample_args)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(warmup):
            out = block(*args)
            outs = out if isinstance(out, (tuple, list)) else (out,)
            grads = [torch.ones_like(o) for o in outs if o.requires_grad]
            need = [a for a in args if a.requires_grad]
            params = [p for p in block.parameters() if p.requires_grad] if isinstance(block, nn.Module) else []
            if grads:
                torch.autograd.grad([o for o in outs if o.requires_grad], need + params, grads, allow_unused=True)
    torch.cuda.current_stream().wait_stream(s)
    return torch.cuda.make_graphed_callables(block, args)


def install_reference_aliases(force: bool = False) -> list[str]:
    """Make the reference's import paths resolve to this library (checkpoints pickle by class path).

    Registers ``models.ops_dcnv3``, ``models.ops_dcnv3.functions``, ``models.ops_dcnv3.functions.dcnv3_func``,
    ``models.ops_dcnv3.modules`` and ``models.ops_dcnv3.modules.dcnv3`` in ``sys.modules`` unless the reference's
    own package is importable (then its files import the top-level ``DCNv3`` shim and nothing is needed) or
    ``force`` is set.  Returns the names it registered.
    """
    pairs = {
        "models.ops_dcnv3": "yolo_somi_b200.ops_dcnv3",
        "models.ops_dcnv3.functions": "yolo_somi_b200.ops_dcnv3.functions",
        "models.ops_dcnv3.functions.dcnv3_func": "yolo_somi_b200.ops_dcnv3.functions.dcnv3_func",
        "models.ops_dcnv3.modules": "yolo_somi_b200.ops_dcnv3.modules",
        "models.ops_dcnv3.modules.dcnv3": "yolo_somi_b200.ops_dcnv3.modules.dcnv3",
    }
    done = []
    for alias, real in pairs.items():
        if alias in sys.modules and not force:
            continue
        if not force:
            try:
                importlib.import_module(alias)     # the reference's own package is on the path
                continue
            except Exception:
                pass
        sys.modules[alias] = importlib.import_module(real)
        done.append(alias)
    if "models" not in sys.modules and done:
        import types
        pkg = types.ModuleType("models")
        pkg.__path__ = []                         # a namespace stub: only the aliases above live in it
        sys.modules["models"] = pkg
    return done


# ----------------------------------------------------------------------------------------------------------------------
# The DDP wiring of the training step (reference train.py:165-208 wrap / SyncBN / sampler shard, :263-283 autocast
# forward - backward - optimizer - EMA - logging, :422-434 process-group set-up), restated for one process per GPU.
#
# What changes against the reference, and why:
#   * the device is bound by LOCAL_RANK only (train.py:55 pins GPU 0 through CUDA_VISIBLE_DEVICES and --device '0' is
#     the default, train.py:374: under torchrun every rank would land on the same GPU);
#   * bf16 autocast without a GradScaler (train.py:263 uses fp16 + scaler; the scaler's inf check is a host sync per step);
#   * DDP with gradient_as_bucket_view (the all-reduce runs on the gradients' own storage, no bucket copy-out),
#     static_graph (no per-iteration graph search) and broadcast_buffers=False (BatchNorm statistics are per rank in
#     the reference as well unless --sync-bn; no per-step broadcast of the buffers);
#   * accumulation micro-steps run under no_sync(): one all-reduce per OPTIMIZER step, not per micro-step (train.py:270
#     reduces on every backward);
#   * the running loss stays on the device; the host reads it every `log_every` steps through a pinned buffer and an
#     event instead of formatting it into tqdm every step (train.py:279-282 synchronises on every iteration);
#   * the EMA is FusedModelEMA (above), on every rank-0 optimizer step.
import contextlib
import os

import torch.distributed as dist


def setup_process_group(backend: str | None = None) -> tuple[int, int, int]:
    """(rank, local_rank, world) from the torchrun environment; initialises the default group if world > 1.
    NCCL on GPUs (over NVLink 5 / NVSwitch), gloo on CPU (tests)."""
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        kw = {}
        if use_cuda and (backend or "nccl") == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"), rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def wrap_ddp(model: nn.Module, local_rank: int | None = None, sync_bn: bool = False, stream=None) -> nn.Module:
    """train.py:165-208: optional SyncBatchNorm conversion, then DistributedDataParallel -- with the settings above.
    Returns the model unchanged when there is one rank.  (static_graph: the FIRST backward must be a synchronised one --
    DDP records the graph there -- so accumulation starts with a closing micro-step, as the reference's warm-up does.)"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return model
    on_gpu = next(model.parameters()).is_cuda
    if sync_bn and on_gpu:
        model = nn.SyncBatchNorm.convert_sync_batchnorm(model)
    kw = dict(device_ids=[local_rank], output_device=local_rank) if on_gpu else {}
    # `stream`: a step that will be captured whole (TrainStep(graph=True, graph_stream=stream)) needs DDP constructed on
    # the side stream the capture will run on (PyTorch's CUDA-graphs note on DistributedDataParallel)
    ctx = contextlib.nullcontext()
    if stream is not None and on_gpu:
        stream.wait_stream(torch.cuda.current_stream())
        ctx = torch.cuda.stream(stream)
    with ctx:
        ddp = nn.parallel.DistributedDataParallel(model, gradient_as_bucket_view=True, static_graph=True,
                                                  broadcast_buffers=False, bucket_cap_mb=64, **kw)
    if stream is not None and on_gpu:
        torch.cuda.current_stream().wait_stream(stream)
    return ddp


def make_optimizer(model: nn.Module, lr: float = 0.01, momentum: float = 0.937, weight_decay: float = 5e-4):
    """The reference's three parameter groups (utils/torch_utils.py smart_optimizer: weights with decay, BatchNorm /
    LayerNorm weights and all biases without), SGD with Nesterov momentum (data/hyps/hyp.VisDrone.yaml), multi-tensor."""
    decay, no_decay = [], []
    norms = tuple(v for k, v in nn.__dict__.items() if "Norm" in k and isinstance(v, type))
    for mod in model.modules():
        for name, p in mod.named_parameters(recurse=False):
            if not p.requires_grad:
                continue
            (no_decay if name == "bias" or isinstance(mod, norms) else decay).append(p)
    groups = [{"params": decay, "weight_decay": weight_decay}, {"params": no_decay, "weight_decay": 0.0}]
    on_gpu = any(p.is_cuda for p in decay + no_decay)
    # fused on the GPU: one multi-tensor kernel, and it takes GradScaler's found_inf tensor, so an fp16 step that
    # overflowed is skipped ON THE DEVICE (the unfused path reads found_inf with .item(): a host sync per step)
    return torch.optim.SGD(groups, lr=lr, momentum=momentum, nesterov=True, **({"fused": True} if on_gpu else {"foreach": True}))


def yolo_surrogate_loss(preds, targets, nc: int):
    """A YOLO-shaped loss for the step harness: objectness BCE over every anchor of every level, class BCE and box L1 at
    the cells the targets fall into.  (The reference's ComputeLoss -- build_targets' anchor matching + CIoU,
    utils/loss.py -- is outside the hot path; this has the same inputs, touches every output element
    and back-propagates through every layer, without a host synchronisation.)
    preds: list of [B, na, H, W, nc + 5]; targets: [T, 6] = (image, class, x, y, w, h) normalised."""
    img, cls = targets[:, 0].long(), targets[:, 1].long()
    total = preds[0].new_zeros((), dtype=torch.float32)
    for p in preds:
        b, na, h, w, no = p.shape
        p = p.float()
        gx = (targets[:, 2] * w).long().clamp_(0, w - 1)
        gy = (targets[:, 3] * h).long().clamp_(0, h - 1)
        obj = torch.zeros(b, na, h, w, device=p.device)
        obj[img, :, gy, gx] = obj.new_ones(())           # a device scalar: no host -> device copy inside the step
        total = total + nn.functional.binary_cross_entropy_with_logits(p[..., 4], obj)
        sel = p[img, :, gy, gx]                                         # [T, na, no]
        want_cls = nn.functional.one_hot(cls, nc).float()[:, None].expand(-1, na, -1)
        total = total + 0.5 * nn.functional.binary_cross_entropy_with_logits(sel[..., 5:], want_cls)
        total = total + 0.05 * (sel[..., :4].sigmoid() - targets[:, None, 2:6]).abs().mean()
    return total


class TrainStep:
    """One optimizer step of the reference's loop (train.py:249-283) without its per-iteration host synchronisations.

    step(imgs, targets, last_micro=True): autocast forward, loss, backward (under no_sync() unless `last_micro`), and on
    the last micro-step optimizer step + zero_grad + EMA.  The running loss sum lives on the device; `loss_for_log()`
    returns the mean recorded at the last `log_every` boundary (read through a pinned buffer guarded by an event).

    graph=True: the WHOLE step -- forward, loss, backward with DDP's bucketed all-reduce, unscale, clip, fused SGD,
    EMA, loss sum -- is captured into one CUDA graph after `graph_after` eager steps (each a real step on its own
    batch, run on the capture stream so that everything created lazily exists: cuDNN plans, tensor maps, DDP's
    rebuilt buckets and its first ten iterations of runtime statistics) and replayed from static input buffers.
    At 16 images per GPU (global batch 128 on eight GPUs) the eager step is bound by the host's enqueue time of
    ~3000 kernels, not by the GPU (profiles/README.md); a replay is one launch.  Whole steps only (no accumulation),
    static shapes; the EMA's decay schedule runs on the device (`FusedModelEMA.update(device_count=...)`).  A learning
    rate schedule must write into a tensor lr (`make_optimizer(..., lr=torch.tensor(...))`): a float is baked in."""

    def __init__(self, model: nn.Module, nc: int, optimizer=None, ema: FusedModelEMA | None = None,
                 autocast_dtype: torch.dtype | None = torch.bfloat16, log_every: int = 50, max_norm: float = 10.0,
                 graph: bool = False, graph_after: int | None = None, graph_stream=None):
        self.model, self.nc = model, nc
        self.optimizer = optimizer or make_optimizer(model)
        self.ema, self.dtype, self.log_every, self.max_norm = ema, autocast_dtype, log_every, max_norm
        dev = next(model.parameters()).device
        self.device = dev
        self._lsum = torch.zeros((), device=dev)
        self._steps = 0
        self._host = torch.zeros((), pin_memory=True) if dev.type == "cuda" else torch.zeros(())
        self._event = torch.cuda.Event() if dev.type == "cuda" else None
        self._logged = None
        # fp16 autocast (the reference's AMP dtype, train.py:263) needs the reference's GradScaler (train.py:217,268-274)
        self.scaler = torch.amp.GradScaler(dev.type) if (autocast_dtype == torch.float16 and dev.type == "cuda") else None
        self.graph = bool(graph) and dev.type == "cuda"
        ddp = isinstance(model, nn.parallel.DistributedDataParallel)
        self.graph_after = (11 if ddp else 3) if graph_after is None else graph_after
        self._g = self._static = self._ema_count = None
        self._stream = graph_stream
        self.graph_error = None

    # -- the step's device work: everything between the inputs and the loss sum -------------------------------------
    def _work(self, imgs: torch.Tensor, targets: torch.Tensor, last_micro: bool, ema_count=None) -> torch.Tensor:
        ddp = isinstance(self.model, nn.parallel.DistributedDataParallel)
        sync_ctx = self.model.no_sync() if (ddp and not last_micro) else contextlib.nullcontext()
        amp = torch.autocast(self.device.type, dtype=self.dtype) if self.dtype is not None else contextlib.nullcontext()
        with sync_ctx:
            with amp:
                preds = self.model(imgs)
            loss = yolo_surrogate_loss(preds, targets, self.nc)
            # train.py:266-267 multiplies the loss by WORLD_SIZE because the reference's loss is a per-rank SUM (utils/loss.py
            # returns loss * batch_size) that DDP's gradient averaging would otherwise shrink.  This harness' loss is a
            # MEAN over the local batch: DDP's average of the ranks' gradients already is the gradient of the global mean,
            # so the update for a given global batch is the same at every N.  (With the factor, eight ranks took 8x larger
            # steps, the learned offsets left the value kernel's band within a dozen steps and the backward fell back to
            # its plane form: profiles/README.md, round 2.)
            (self.scaler.scale(loss) if self.scaler is not None else loss).backward()
        if last_micro:
            if self.scaler is not None:
                self.scaler.unscale_(self.optimizer)                                                   # train.py:271
            if self.max_norm:
                torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.max_norm, foreach=True)   # train.py:272
            if self.scaler is not None:
                self.scaler.step(self.optimizer)      # fused optimizer: found_inf stays on the device
                self.scaler.update()
            else:
                self.optimizer.step()
            self.optimizer.zero_grad(set_to_none=True)
            if self.ema is not None:
                self.ema.update(self.model, device_count=ema_count)
            self._lsum += loss.detach()
        return loss.detach()

    def _after(self) -> None:
        self._steps += 1
        if self._steps % self.log_every == 0:
            self._host.copy_(self._lsum, non_blocking=True)
            if self._event is not None:
                self._event.record()
            self._logged = self._steps

    def step(self, imgs: torch.Tensor, targets: torch.Tensor, last_micro: bool = True) -> torch.Tensor:
        if self.graph and self.graph_error is None:
            if not last_micro:
                raise ValueError("TrainStep(graph=True) captures whole optimizer steps; accumulate with graph=False")
            return self._graph_step(imgs, targets)
        loss = self._work(imgs, targets, last_micro)
        if last_micro:
            self._after()
        return loss

    # -- graph mode -------------------------------------------------------------------------------------------------
    def _graph_step(self, imgs: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        cur = torch.cuda.current_stream(self.device)
        if self._static is None:
            if self._stream is None:
                self._stream = torch.cuda.Stream(self.device)
            self._static = (torch.empty_like(imgs), torch.empty_like(targets))       # preserves channels_last strides
            self._loss_out = torch.zeros((), device=self.device)
            if self.ema is not None:
                self._ema_count = torch.full((), float(self.ema.updates), dtype=torch.float64, device=self.device)
        sx, st = self._static
        if sx.shape != imgs.shape or st.shape != targets.shape:
            raise ValueError("TrainStep(graph=True): static shapes only (drop the last, ragged batch)")
        # the copies into the static buffers run on the caller's stream; the step's stream waits for them
        sx.copy_(imgs, non_blocking=True)
        st.copy_(targets, non_blocking=True)
        self._stream.wait_stream(cur)
        with torch.cuda.stream(self._stream):
            if self._g is None and self._steps >= self.graph_after:
                try:
                    self._capture()
                except Exception as exc:                  # stay correct: fall back to eager steps, say why
                    self.graph_error = "%s: %s" % (type(exc).__name__, str(exc)[:300])
                    self._g = None
            if self._g is not None:
                self._g.replay()
                if self.ema is not None:
                    self.ema.updates += 1
            else:
                # eager warm-up steps (real steps) -- and the fall-back if capture failed
                self._loss_out.copy_(self._work(sx, st, True, ema_count=self._ema_count if self.graph_error is None else None))
                if self.ema is not None and self.graph_error is None:
                    self.ema.updates += 1
            self._after()
        cur.wait_stream(self._stream)
        return self._loss_out

    def _capture(self) -> None:
        sx, st = self._static
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(g, stream=self._stream):
            self._loss_out.copy_(self._work(sx, st, True, ema_count=self._ema_count))
        self._g = g

    def loss_for_log(self):
        """(step, mean loss) of the last logging boundary, or None; waits only for that boundary's copy."""
        if self._logged is None:
            return None
        if self._event is not None:
            self._event.synchronize()
        return self._logged, float(self._host) / self._logged


def synthetic_batch(batch: int, size: int = 640, nc: int = 10, boxes_per_image: int = 64, device="cuda", seed: int = 0):
    """VisDrone-shaped synthetic batch: imgs ~ U[0, 1) [B, 3, S, S]; targets [B * 64, 6] = (image, class,
    x, y, w, h) with tiny boxes (w, h ~ logU(0.005, 0.08))."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    imgs = torch.rand(batch, 3, size, size, generator=g)
    t = batch * boxes_per_image
    wh = torch.exp(torch.empty(t, 2).uniform_(math.log(0.005), math.log(0.08), generator=g))
    targets = torch.cat([torch.arange(batch).repeat_interleave(boxes_per_image)[:, None].float(),
                         torch.randint(0, nc, (t, 1), generator=g).float(), torch.rand(t, 2, generator=g), wh], 1)
    return imgs.to(device), targets.to(device)
