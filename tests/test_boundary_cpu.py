"""CPU (no GPU needed): the drop-in boundary -- C ABI exports, error codes, shim validation,
signature compatibility with the reference's own call sites, sharding helpers under gloo."""
import ctypes
import inspect
import os
import re
import sys
import types
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from yolo_somi_b200 import build, _native
    build.build()
    return _native.load()


def test_library_exports_every_declared_symbol(lib):
    header = (ROOT / "include" / "dcnv3_sm100.h").read_text()
    declared = set(re.findall(r"DCNV3_API\s+[\w\s\*]+?\b(dcnv3_\w+)\s*\(", header))
    from yolo_somi_b200 import _native
    assert declared == set(_native.EXPORTS), (declared, _native.EXPORTS)
    raw = ctypes.CDLL(str(_native.LIB_PATH))
    for sym in declared:
        assert getattr(raw, sym) is not None
    assert lib.dcnv3_sm100_abi_version() == 1


def test_error_codes_without_touching_a_gpu(lib):
    """Argument validation happens before any CUDA call, so it is testable here."""
    from yolo_somi_b200 import _native
    geom_ok = (1, 8, 8, 8, 8, 2, 4, 3, 3, 1, 1, 1, 1, 1, 1)   # N H W Ho Wo G gc kh kw sh sw ph pw dh dw
    call = lambda ptrs, geom, dtype: lib.dcnv3_forward_sm100(*ptrs, *geom, 1.0, dtype, None)
    assert call((None,) * 4, geom_ok, 0) == -3                          # DCNV3_E_NULL
    assert call((None,) * 4, geom_ok, 7) == -1                          # DCNV3_E_DTYPE
    bad = list(geom_ok); bad[3] = 9                                     # Ho inconsistent
    assert call((None,) * 4, bad, 0) == -2                              # DCNV3_E_SHAPE
    zero = list(geom_ok); zero[0] = 0
    assert call((None,) * 4, zero, 0) == 0                              # empty batch is a no-op
    big = (1, 70000, 70000, 70000, 70000, 2, 4, 3, 3, 1, 1, 1, 1, 1, 1)
    assert call((None,) * 4, big, 0) == -5                              # DCNV3_E_TOO_LARGE
    assert call((16, 2, 16, 16), geom_ok, 2) == -6                      # offset misaligned (bf16 pair)
    assert "dtype" in _native.strerror(-1) and _native.strerror(0) == "ok"
    # workspace sizing: none for fp32, fp32 plane (+header) for 16-bit, int64 plane when deterministic
    ws = lib.dcnv3_backward_workspace_bytes
    plane = 2 * 8 * 8 * 8
    assert ws(2, 8, 8, 2, 4, 0, 0) == 0
    assert ws(2, 8, 8, 2, 4, 2, 0) == 256 + plane * 4
    assert ws(2, 8, 8, 2, 4, 0, 1) == 256 + plane * 8
    rc = lib.dcnv3_backward_sm100(16, 16, 16, 16, 16, 16, 16, 16, 8, *geom_ok, 1.0, 2, 0, None)
    assert rc == -4                                                     # DCNV3_E_WORKSPACE
    # fp64 (DCNV3_F64 = 3, the dtype the reference's own test script starts with): a known dtype, no workspace,
    # 16-byte (dx, dy) pairs, and no deterministic mode
    assert _native.F64 == 3 and call((None,) * 4, geom_ok, 3) == -3     # DCNV3_E_NULL, not DCNV3_E_DTYPE
    assert call((16, 8, 16, 16), geom_ok, 3) == -6                      # offset not aligned to a double pair
    assert ws(2, 8, 8, 2, 4, 3, 0) == 0
    rc = lib.dcnv3_backward_sm100(16, 16, 16, 16, 16, 16, 16, None, 0, *geom_ok, 1.0, 3, 1, None)
    assert rc == -1 and "fp64" in _native.strerror(-1)                  # DETERMINISTIC | fp64


def test_shim_rejects_what_the_reference_rejects():
    """Same RuntimeErrors as dcnv3_cuda.cu:29-53 / dcnv3.h:37 (no device needed for these)."""
    import DCNv3
    v, o, m = torch.zeros(2, 4, 4, 8), torch.zeros(2, 4, 4, 36), torch.zeros(2, 4, 4, 18)
    geom = (3, 3, 1, 1, 1, 1, 1, 1, 2, 4, 1.0)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        DCNv3.dcnv3_forward(v, o, m, *geom, 256)
    with pytest.raises(RuntimeError, match="contiguous"):
        DCNv3.dcnv3_forward(v.transpose(1, 2), o, m, *geom, 256)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        DCNv3.dcnv3_backward(v, o, m, *geom, v.clone(), 256)


def test_signatures_match_the_reference_call_sites():
    """Positional order of src/dcnv3.h:20-59 (grad_output sits BEFORE im2col_step)."""
    import DCNv3
    fwd = list(inspect.signature(DCNv3.dcnv3_forward).parameters)
    bwd = list(inspect.signature(DCNv3.dcnv3_backward).parameters)
    common = ["input", "offset", "mask", "kernel_h", "kernel_w", "stride_h", "stride_w", "pad_h",
              "pad_w", "dilation_h", "dilation_w", "group", "group_channels", "offset_scale"]
    assert fwd == common + ["im2col_step"]
    assert bwd == common + ["grad_output", "im2col_step"]
    from yolo_somi_b200.ops_dcnv3.functions import DCNv3Function
    assert list(inspect.signature(DCNv3Function.forward).parameters)[1:] == common + ["im2col_step"]
    from yolo_somi_b200.ops_dcnv3.modules import DCNv3 as Layer
    ref_ctor = ["channels", "kernel_size", "dw_kernel_size", "stride", "pad", "dilation", "group",
                "offset_scale", "act_layer", "norm_layer", "center_feature_scale", "use_dcn_v4_op"]
    assert list(inspect.signature(Layer.__init__).parameters)[1:] == ref_ctor


@pytest.mark.skipif(not Path("/root/reference/models/ops_dcnv3").exists(),
                    reason="reference tree only exists in the build container")
def test_reference_python_files_import_unmodified_on_top_of_the_shim():
    """`import DCNv3` (functions/dcnv3_func.py:16) resolves to this repo's module, and the
    reference's own DCNv3Function / DCNv3 layer construct against it."""
    import DCNv3  # noqa: F401  (repo root is on sys.path)
    sys.path.insert(0, "/root/reference")
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "models" or k.startswith("models.")}
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            from models.ops_dcnv3.functions.dcnv3_func import DCNv3Function as RefFn
            from models.ops_dcnv3.modules.dcnv3 import DCNv3 as RefLayer
        assert sys.modules["DCNv3"].__file__.startswith(str(ROOT))
        layer = RefLayer(channels=32, group=2)
        from yolo_somi_b200.ops_dcnv3.modules import DCNv3 as OurLayer
        assert sorted(layer.state_dict()) == sorted(OurLayer(channels=32, group=2).state_dict())
        with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
            RefFn.apply(torch.zeros(1, 4, 4, 8), torch.zeros(1, 4, 4, 36), torch.zeros(1, 4, 4, 18),
                        3, 3, 1, 1, 1, 1, 1, 1, 2, 4, 1.0, 256)
    finally:
        sys.path.remove("/root/reference")
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from yolo_somi_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(_native.DCNv3NativeError, match="no CPU or PyTorch fallback"):
        _native.load()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under yolo_somi_b200/ or DCNv3.py may reference it."""
    offenders = []
    for path in list((ROOT / "yolo_somi_b200").rglob("*.py")) + [ROOT / "DCNv3.py"]:
        text = path.read_text()
        if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M):
            offenders.append(str(path))
    assert not offenders, offenders


# ----------------------------------------------------------------------------- sharding (gloo)
def test_shard_range_partitions_the_batch():
    from yolo_somi_b200.sharding import shard_range
    for n, world in ((16, 1), (16, 2), (16, 8), (17, 4), (3, 8), (0, 2)):
        spans = [shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _gloo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from yolo_somi_b200.sharding import shard_range, whole_job_throughput, max_over_ranks
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = shard_range(17, rank, world)
        # rank r "processes" its images in (r+1) seconds: the job rate is total / slowest
        rate = whole_job_throughput(float(e - b), float(rank + 1))
        slowest = max_over_ranks(float(rank + 1))
        out[rank] = (b, e, rate, slowest)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_aggregation():
    import torch.multiprocessing as mp
    world, port = 2, 29600 + os.getpid() % 300
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_gloo_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert res[0][:2] == (0, 9) and res[1][:2] == (9, 17)
    for r in range(world):
        assert res[r][2] == pytest.approx(17 / 2.0) and res[r][3] == 2.0


# ----------------------------------------------------------------------------- ONNX symbolic
def test_symbolic_emits_the_reference_node():
    """DCNv3Function.symbolic (reference dcnv3_func.py:63-89): node `mmdeploy::TRTDCNv3`, three tensor inputs, eleven
    integer attributes (`_i`) and `offset_scale_f`, exactly the reference's names."""
    from yolo_somi_b200.ops_dcnv3.functions import DCNv3Function

    class G:
        def op(self, name, *inputs, **attrs):
            self.call = (name, inputs, attrs)
            return "node"
    g = G()
    out = DCNv3Function.symbolic(g, "x", "off", "msk", 3, 3, 1, 1, 1, 1, 1, 1, 4, 16, 2.0, 256)
    name, inputs, attrs = g.call
    assert out == "node" and name == "mmdeploy::TRTDCNv3" and inputs == ("x", "off", "msk")
    assert attrs == dict(kernel_h_i=3, kernel_w_i=3, stride_h_i=1, stride_w_i=1, pad_h_i=1, pad_w_i=1,
                         dilation_h_i=1, dilation_w_i=1, group_i=4, group_channels_i=16, offset_scale_f=2.0,
                         im2col_step_i=256)
    from baseline import ref_loader
    if ref_loader.available():          # the reference's own symbolic, fed the same arguments
        RefFn, _, _ = ref_loader.dropin()
        g2 = G()
        RefFn.symbolic(g2, "x", "off", "msk", 3, 3, 1, 1, 1, 1, 1, 1, 4, 16, 2.0, 256)
        assert g2.call == g.call
