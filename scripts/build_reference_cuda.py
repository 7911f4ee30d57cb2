#!/usr/bin/env python
"""Build the REFERENCE's own CUDA extension for sm_100a as a timing comparator (last row).

    python scripts/build_reference_cuda.py            # -> baseline/_ref/DCNv3_refcuda.so (git-ignored)

The reference's `models/ops_dcnv3/src` does not compile against torch 2.11 as it stands: the two AT_DISPATCH calls in
`src/cuda/dcnv3_cuda.cu:70,148` pass `input.type()` (an `at::DeprecatedTypeProperties`) where a `c10::ScalarType` is
required.  This script copies the staged sources (`scripts/stage_reference.py`) to a scratch directory, replaces
those two tokens by `input.scalar_type()` -- nothing else -- and compiles the extension under the module name
`DCNv3_refcuda`, so it can never be mistaken for this repo's `DCNv3` module.  The result is "the reference's kernels
recompiled for B200": `scripts/bench_reference_cuda.py` and `bench.py`'s `reference_cuda_kernels` row time it next to
this library on the same box (fp16: the reference has no bf16 dispatch).  It is not product source, nothing under
`yolo_somi_b200/` loads it, and no source of it enters the repository.
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "baseline" / "_ref"
SRC = REF / "models" / "ops_dcnv3" / "src"
OUT = REF / "DCNv3_refcuda.so"


def patch_dispatch(text: str) -> str:
    """The whole patch: `input.type()` -> `input.scalar_type()` as the first argument of the two AT_DISPATCH calls."""
    patched = text.replace("input.type(), \"ms_deform_attn_forward_cuda\"", "input.scalar_type(), \"ms_deform_attn_forward_cuda\"")
    patched = patched.replace("input.type(), \"ms_deform_attn_backward_cuda\"", "input.scalar_type(), \"ms_deform_attn_backward_cuda\"")
    if patched.count("input.scalar_type(), ") != 2:
        raise RuntimeError("expected exactly two AT_DISPATCH sites to patch in dcnv3_cuda.cu")
    return patched


def build(verbose: bool = False) -> Path:
    if not SRC.is_dir():
        raise FileNotFoundError(f"{SRC}: run scripts/stage_reference.py first")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils import cpp_extension

    work = Path(tempfile.mkdtemp(prefix="dcnv3_refcuda_"))
    try:
        src = work / "src"
        shutil.copytree(SRC, src)
        cu = src / "cuda" / "dcnv3_cuda.cu"
        text = cu.read_text()
        patched = patch_dispatch(text)
        cu.write_text(patched)
        sources = [str(src / "vision.cpp"), str(src / "cpu" / "dcnv3_cpu.cpp"), str(cu)]
        bdir = work / "build"
        bdir.mkdir()
        cpp_extension.load(
            name="DCNv3_refcuda", sources=sources, extra_include_paths=[str(src)],
            extra_cflags=["-DWITH_CUDA", "-O3"],
            extra_cuda_cflags=["-DWITH_CUDA", "-O3", "-gencode", "arch=compute_100a,code=sm_100a"],
            build_directory=str(bdir), verbose=verbose, is_python_module=False)
        shutil.copy2(bdir / "DCNv3_refcuda.so", OUT)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    return OUT


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print(f"built {p} ({p.stat().st_size >> 10} KiB)")
