#!/usr/bin/env python
"""Stage the reference's own DCNv3 package, verbatim, under git-ignored baseline/_ref/ (SURVEY F10).

    python scripts/stage_reference.py [--src /root/reference]

Copies <src>/models/ops_dcnv3 (Python files, test.py, the C++/CUDA sources for the record) to
baseline/_ref/models/ops_dcnv3 byte for byte and writes baseline/_ref/MANIFEST.json (sha256 per file).
The copy is NOT product source and never enters git history (.gitignore: baseline/_ref/); it travels to the
GPU box with the snapshot, where /root/reference does not exist.  Users:
  * bench.py --impl reference / the cpu_baseline leg: the unmodified dcnv3_core_pytorch on the host cores;
  * tests/test_reference_dropin.py: the reference's own DCNv3Function / DCNv3 layer running on the GPU over
    this repo's `DCNv3` module.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
DST = ROOT / "baseline" / "_ref"


def stage(src: Path) -> dict:
    pkg = src / "models" / "ops_dcnv3"
    if not pkg.is_dir():
        raise FileNotFoundError(pkg)
    out = DST / "models" / "ops_dcnv3"
    if out.exists():
        shutil.rmtree(out)
    shutil.copytree(pkg, out, ignore=shutil.ignore_patterns("__pycache__", "*.so", "build", "*.egg-info"))
    manifest = {}
    for p in sorted(out.rglob("*")):
        if p.is_file():
            manifest[str(p.relative_to(DST))] = hashlib.sha256(p.read_bytes()).hexdigest()
    (DST / "MANIFEST.json").write_text(json.dumps({"source": str(pkg), "files": manifest}, indent=1))
    return manifest


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    a = ap.parse_args()
    m = stage(Path(a.src))
    print(f"staged {len(m)} files under {DST}")
    sys.exit(0)
