"""Import the UNMODIFIED reference package staged under baseline/_ref/ (scripts/stage_reference.py).

Two ways in, matching the two things the staged copy is for:
  core_pytorch()  -> the reference's CPU path `dcnv3_core_pytorch` (functions/dcnv3_func.py:147-188) with the compiled
                     module it never calls stubbed out: the reference arm of bench.py and the cpu_baseline leg;
  dropin()        -> the reference's own `DCNv3Function`, `DCNv3` layer and `DCNv3_pytorch` with `import DCNv3`
                     (functions/dcnv3_func.py:16) resolving to THIS repo's module: the drop-in claim, run on a GPU.
Nothing under yolo_somi_b200/ imports this file.
"""
from __future__ import annotations

import importlib
import sys
import types
import warnings
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "baseline" / "_ref"


def available() -> bool:
    return (REF / "models" / "ops_dcnv3" / "functions" / "dcnv3_func.py").is_file()


def _purge():
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]


def _import(names):
    if not available():
        raise FileNotFoundError(f"{REF}: run scripts/stage_reference.py where /root/reference exists")
    _purge()
    sys.path.insert(0, str(REF))
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")          # torch.cuda.amp.custom_fwd deprecation (dcnv3_func.py:15)
            out = []
            for mod, attr in names:
                out.append(getattr(importlib.import_module(mod), attr))
        return out
    finally:
        sys.path.remove(str(REF))


def core_pytorch():
    """The reference's dcnv3_core_pytorch, importable without any compiled extension."""
    stub = "DCNv3" not in sys.modules
    if stub:
        sys.modules["DCNv3"] = types.ModuleType("DCNv3")     # imported at dcnv3_func.py:16, never called on this path
    try:
        (fn,) = _import([("models.ops_dcnv3.functions.dcnv3_func", "dcnv3_core_pytorch")])
    finally:
        if stub:
            del sys.modules["DCNv3"]
        _purge()
    return fn


def dropin():
    """(DCNv3Function, DCNv3 layer, DCNv3_pytorch layer) of the reference, bound to this repo's `DCNv3` module."""
    if str(ROOT) not in sys.path:
        sys.path.insert(0, str(ROOT))
    import DCNv3  # noqa: F401  the repo-root shim
    assert Path(sys.modules["DCNv3"].__file__).resolve().parent == ROOT
    return _import([("models.ops_dcnv3.functions.dcnv3_func", "DCNv3Function"),
                    ("models.ops_dcnv3.modules.dcnv3", "DCNv3"),
                    ("models.ops_dcnv3.modules.dcnv3", "DCNv3_pytorch")])
