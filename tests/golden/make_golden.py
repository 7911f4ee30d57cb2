"""Generate tests/golden/*.npz by running the REAL reference implementation.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports ``dcnv3_core_pytorch`` (models/ops_dcnv3/functions/dcnv3_func.py:147-188) and
``DCNv3_pytorch`` (models/ops_dcnv3/modules/dcnv3.py:95-219) from the read-only reference tree.
``functions/dcnv3_func.py:16`` does ``import DCNv3`` (the compiled extension) at import time; the
oracle path never calls it, so an empty stand-in module is registered first.  Nothing is copied
from the reference: only its numerical outputs are stored.
"""
from __future__ import annotations

import sys
import types
import warnings
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import cases  # noqa: E402

REF = Path("/root/reference")


def load_reference():
    if not REF.exists():
        raise SystemExit("needs /root/reference (build container only)")
    sys.modules.setdefault("DCNv3", types.ModuleType("DCNv3"))
    sys.path.insert(0, str(REF))
    warnings.filterwarnings("ignore")
    from models.ops_dcnv3.functions.dcnv3_func import dcnv3_core_pytorch
    from models.ops_dcnv3.modules.dcnv3 import DCNv3_pytorch
    return dcnv3_core_pytorch, DCNv3_pytorch


def store(bag, key, arr):
    a = np.asarray(arr)
    bag[key + "/sum"] = np.float64(a.astype(np.float64).sum())
    if a.size > cases.BIG:
        bag[key + "/sampled"] = a.reshape(-1)[::cases.SAMPLE_STRIDE].copy()
    else:
        bag[key + "/full"] = a


def main():
    core, module_cls = load_reference()
    torch.set_num_threads(8)
    bag = {}
    for c in cases.ALL:
        arrs = cases.make_inputs(c)
        bag[f"{c.name}/insum"] = np.float64(cases.input_checksum(arrs))
        for dt, tdt in (("f64", torch.float64), ("f32", torch.float32)):
            v, o, m, g = (torch.from_numpy(a).to(tdt) for a in arrs)
            v.requires_grad_(True); o.requires_grad_(True); m.requires_grad_(True)
            out = core(v, o, m, *c.geom)
            out.backward(g)
            store(bag, f"{c.name}/{dt}/out", out.detach().numpy())
            store(bag, f"{c.name}/{dt}/gv", v.grad.numpy())
            store(bag, f"{c.name}/{dt}/go", o.grad.numpy())
            store(bag, f"{c.name}/{dt}/gm", m.grad.numpy())
        print("core", c.name, "ok")
    np.savez_compressed(HERE / "core.npz", **bag)

    bag = {}
    for mc in cases.MODULE_CASES:
        state, x, grad = cases.make_module_state(mc)
        mod = module_cls(channels=mc.channels, kernel_size=mc.kernel_size, stride=mc.stride,
                         pad=mc.pad, dilation=mc.dilation, group=mc.group,
                         offset_scale=mc.offset_scale,
                         center_feature_scale=mc.center_feature_scale)
        missing = mod.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()},
                                      strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        xin = torch.from_numpy(x).requires_grad_(True)
        y = mod(xin)
        y.backward(torch.from_numpy(grad))
        bag[f"{mc.name}/insum"] = np.float64(
            cases.input_checksum([x, grad] + [state[k] for k in sorted(state)]))
        bag[f"{mc.name}/keys"] = np.array(sorted(mod.state_dict().keys()))
        bag[f"{mc.name}/y"] = y.detach().numpy()
        bag[f"{mc.name}/gx"] = xin.grad.numpy()
        for k, p in mod.named_parameters():
            bag[f"{mc.name}/gp/{k}"] = p.grad.numpy()
        print("module", mc.name, "ok")
    np.savez_compressed(HERE / "module.npz", **bag)
    for f in ("core.npz", "module.npz"):
        print(f, (HERE / f).stat().st_size, "bytes")


if __name__ == "__main__":
    main()
