"""ctypes binding of libdcnv3_sm100.so (the C ABI in include/dcnv3_sm100.h).

There is deliberately NO fallback: if the library is missing or a launch fails this raises.
Nothing here imports ``oracle/`` (the oracle is test infrastructure only).
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("DCNV3_SM100_LIB", _PKG / "libdcnv3_sm100.so"))

ABI_VERSION = 1
F32, F16, BF16, F64 = 0, 1, 2, 3
BWD_DETERMINISTIC = 1

EXPORTS = ("dcnv3_sm100_abi_version", "dcnv3_sm100_strerror", "dcnv3_forward_sm100",
           "dcnv3_backward_workspace_bytes", "dcnv3_backward_sm100",
           "dcnv3_host_pipeline_create", "dcnv3_host_pipeline_run", "dcnv3_host_pipeline_sync",
           "dcnv3_host_pipeline_destroy", "dcnv3_offset_mask_proj_padded_cols",
           "dcnv3_offset_mask_proj_sm100", "dcnv3_dwconv_ln_gelu_sm100",
           "dcnv3_dwconv_ln_gelu_backward_sm100", "dcnv3_mask_softmax_backward_sm100", "dcnv3_bias_act_sm100")

_lib = None


class DCNv3NativeError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load (once) and type the library; raises if it is absent or has the wrong ABI."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise DCNv3NativeError(
            f"{LIB_PATH} not found: build it with `python -m yolo_somi_b200.build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the DCNv3 core.")
    lib = ctypes.CDLL(str(LIB_PATH))
    c_int, c_vp, c_f, c_sz, c_u = (ctypes.c_int, ctypes.c_void_p, ctypes.c_float, ctypes.c_size_t,
                                   ctypes.c_uint)
    lib.dcnv3_sm100_abi_version.restype = c_int
    lib.dcnv3_sm100_abi_version.argtypes = []
    lib.dcnv3_sm100_strerror.restype = ctypes.c_char_p
    lib.dcnv3_sm100_strerror.argtypes = [c_int]
    geom = [c_int] * 15  # N H W Ho Wo G gc kh kw sh sw ph pw dh dw
    lib.dcnv3_forward_sm100.restype = c_int
    lib.dcnv3_forward_sm100.argtypes = [c_vp] * 4 + geom + [c_f, c_int, c_vp]
    lib.dcnv3_backward_workspace_bytes.restype = c_sz
    lib.dcnv3_backward_workspace_bytes.argtypes = [c_int] * 6 + [c_u]
    lib.dcnv3_backward_sm100.restype = c_int
    lib.dcnv3_backward_sm100.argtypes = [c_vp] * 8 + [c_sz] + geom + [c_f, c_int, c_u, c_vp]
    lib.dcnv3_host_pipeline_create.restype = c_int
    lib.dcnv3_host_pipeline_create.argtypes = [ctypes.POINTER(c_vp)] + [c_int] * 13 + [c_f, c_int, c_u]
    lib.dcnv3_host_pipeline_run.restype = c_int
    lib.dcnv3_host_pipeline_run.argtypes = [c_vp] * 9 + [c_int]
    lib.dcnv3_host_pipeline_sync.restype = c_int
    lib.dcnv3_host_pipeline_sync.argtypes = [c_vp]
    lib.dcnv3_host_pipeline_destroy.restype = None
    lib.dcnv3_host_pipeline_destroy.argtypes = [c_vp]
    lib.dcnv3_offset_mask_proj_padded_cols.restype = c_int
    lib.dcnv3_offset_mask_proj_padded_cols.argtypes = [c_int, c_int]
    lib.dcnv3_offset_mask_proj_sm100.restype = c_int
    lib.dcnv3_offset_mask_proj_sm100.argtypes = [c_vp] * 5 + [ctypes.c_longlong] + [c_int] * 4 + [c_vp]
    lib.dcnv3_dwconv_ln_gelu_sm100.restype = c_int
    lib.dcnv3_dwconv_ln_gelu_sm100.argtypes = [c_vp] * 7 + [c_int] * 5 + [c_f, c_int, c_vp]
    lib.dcnv3_dwconv_ln_gelu_backward_sm100.restype = c_int
    lib.dcnv3_dwconv_ln_gelu_backward_sm100.argtypes = [c_vp] * 9 + [c_int] * 5 + [c_f, c_int, c_vp]
    lib.dcnv3_mask_softmax_backward_sm100.restype = c_int
    lib.dcnv3_mask_softmax_backward_sm100.argtypes = [c_vp] * 3 + [ctypes.c_longlong, c_int, c_int, c_vp]
    lib.dcnv3_bias_act_sm100.restype = c_int
    lib.dcnv3_bias_act_sm100.argtypes = [c_vp] * 3 + [ctypes.c_longlong, c_int, c_int, c_int, c_vp]
    got = lib.dcnv3_sm100_abi_version()
    if got != ABI_VERSION:
        raise DCNv3NativeError(f"{LIB_PATH}: ABI version {got}, expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def strerror(code: int) -> str:
    return load().dcnv3_sm100_strerror(int(code)).decode()


def check(code: int, what: str) -> None:
    if code != 0:
        raise DCNv3NativeError(f"{what} failed ({code}): {strerror(code)}")
