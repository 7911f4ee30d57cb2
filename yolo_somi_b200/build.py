"""Build recipe for libdcnv3_sm100.so (in-tree, sm_100a only).

    python -m yolo_somi_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
repo snapshot; `__graft_entry__.build()` calls this.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIB = PKG / "libdcnv3_sm100.so"
OBJ_DIR = PKG / "csrc" / "_obj"

SOURCES = ["dcnv3_forward.cu", "dcnv3_forward_tile.cu", "dcnv3_forward_gs.cu", "dcnv3_backward.cu", "dcnv3_backward_mma.cu",
           "dcnv3_backward_strip.cu", "dcnv3_backward_dots.cu", "dcnv3_backward_vmma.cu", "dcnv3_backward_vres.cu", "dcnv3_host_pipeline.cu",
           "dcnv3_f64.cu", "dcnv3_proj.cu", "dcnv3_dwconv.cu", "dcnv3_dwconv_bwd.cu", "dcnv3_hosting.cu", "dcnv3_capi.cu"]
# measured alternatives that lost to the defaults (profiles/README.md): built only on request
EXPERIMENTS = ["experiments/dcnv3_forward_mma.cu", "experiments/dcnv3_backward_mma2.cu", "experiments/dcnv3_backward_tile.cu",
               "experiments/dcnv3_backward_vband.cu"]
WITH_EXPERIMENTS = os.environ.get("DCNV3_BUILD_EXPERIMENTS", "0") not in ("", "0")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",   # no --use_fast_math: IEEE div/sqrt, denormals kept
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _deps() -> list[Path]:
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("experiments/*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + \
        sorted(INCLUDE.glob("*.h"))


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in _deps())


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB
    nvcc = _nvcc()
    OBJ_DIR.mkdir(exist_ok=True)

    sources = SOURCES + (EXPERIMENTS if WITH_EXPERIMENTS else [])
    # (development: DCNV3_NVCC_EXTRA="-DVRES_TPP=2 -DVRES_GROUPS=4" builds a kernel variant)
    flags = NVCC_FLAGS + (["-DDCNV3_EXPERIMENTS"] if WITH_EXPERIMENTS else []) + os.environ.get("DCNV3_NVCC_EXTRA", "").split()

    def compile_one(src: str) -> Path:
        obj = OBJ_DIR / (Path(src).stem + ".o")
        cmd = [nvcc, *flags, "-Xcompiler", "-fPIC,-fvisibility=hidden", f"-I{INCLUDE}", f"-I{CSRC}",
               "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(sources)) as ex:
        objs = list(ex.map(compile_one, sources))
    tmp = LIB.with_suffix(".so.tmp")
    r = subprocess.run([nvcc, "-shared", "-o", str(tmp), *map(str, objs)],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
