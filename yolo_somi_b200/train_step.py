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

    def update(self, model: nn.Module) -> None:
        model = _de_parallel(model)
        with torch.no_grad():
            self.updates += 1
            d = self.decay(self.updates)
            # the tensor lists are rebuilt only when a tensor of either model moved (.to(), new parameters, ...)
            ptrs = self._fingerprint(model, self.ema)
            bound = getattr(self, "_bound", None)
            if bound is None or bound[0] is not model or bound[1] != ptrs:
                self._bind(model, ptrs)
            for es, ss in self._groups:
                torch._foreach_mul_(es, d)                       # v *= d
                scaled = torch._foreach_mul(ss, 1 - d)           # (1 - d) * m   (in m's dtype, as the reference)
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
