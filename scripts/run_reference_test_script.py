#!/usr/bin/env python
"""Run the reference's OWN test script (models/ops_dcnv3/test.py, unmodified, from the copy staged under git-ignored
baseline/_ref/) on a GPU with `import DCNv3` resolving to

    --backend ours      this repository's module (DCNv3.py -> libdcnv3_sm100.so), the drop-in claim end to end
    --backend refcuda   the reference's own CUDA extension recompiled for sm_100a (scripts/build_reference_cuda.py)

and print its output.  The script checks the extension against dcnv3_core_pytorch in double and float (forward, and all
three gradients for group_channels in {1, 16, 30, 32, 64, 71, 1025}: test.py:33-216,257-260) and then times 100 forward
calls at N = 512, 64 x 64, C = 64, G = 4 in fp32 for three im2col_step values (test.py:219-251).
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF_PKG = ROOT / "baseline" / "_ref" / "models" / "ops_dcnv3"
REF_SO = ROOT / "baseline" / "_ref" / "DCNv3_refcuda.so"

_ALIAS = '''import importlib.util as _u, torch as _t
_s = _u.spec_from_file_location("DCNv3_refcuda", %r)
_m = _u.module_from_spec(_s); _s.loader.exec_module(_m)
dcnv3_forward, dcnv3_backward = _m.dcnv3_forward, _m.dcnv3_backward
'''


def run(backend: str = "ours", timeout: int = 900):
    """-> (returncode, stdout, stderr) of `python test.py` in the staged package directory."""
    if not (REF_PKG / "test.py").is_file():
        raise FileNotFoundError(f"{REF_PKG}/test.py: run scripts/stage_reference.py where /root/reference exists")
    env = dict(os.environ)
    with tempfile.TemporaryDirectory(prefix="dcnv3_alias_") as tmp:
        if backend == "ours":
            env["PYTHONPATH"] = str(ROOT) + os.pathsep + env.get("PYTHONPATH", "")
        else:
            if not REF_SO.exists():
                raise FileNotFoundError(f"{REF_SO}: run scripts/build_reference_cuda.py")
            (Path(tmp) / "DCNv3.py").write_text(_ALIAS % str(REF_SO))
            env["PYTHONPATH"] = tmp + os.pathsep + env.get("PYTHONPATH", "")
        r = subprocess.run([sys.executable, "-W", "ignore", "test.py"], cwd=REF_PKG, env=env, capture_output=True, text=True,
                           timeout=timeout)
    return r.returncode, r.stdout, r.stderr


def checks(stdout: str):
    """The script's verdict lines: `* True|False <check>: max_abs_err ... max_rel_err ...`."""
    return [ln for ln in stdout.splitlines() if ln.startswith("* ")]


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="ours", choices=["ours", "refcuda"])
    a = ap.parse_args()
    rc, out, err = run(a.backend)
    print(out, end="")
    if rc:
        print(err[-3000:], file=sys.stderr)
    lines = checks(out)
    print(f"# backend={a.backend} rc={rc} checks={len(lines)} passed={sum(ln.startswith('* True') for ln in lines)}")
    sys.exit(rc)
