"""Host-side logic of the NCHW hosting modules and the parse_model patch (no GPU: constructors and text only)."""
import ast
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_groups_for_always_divides_and_prefers_fast_shapes():
    from yolo_somi_b200.hosting import groups_for
    for c in list(range(1, 300)) + [512, 640, 768, 1024]:
        g = groups_for(c)
        assert g >= 1 and c % g == 0, c
    assert groups_for(256) == 16 and groups_for(128) == 8 and groups_for(512) == 32     # 16 channels per group, G % 8 == 0
    assert groups_for(64) == 8                                                            # 8 per group rather than G = 4
    assert groups_for(100) == 4                                                           # 25 per group: the largest divisor <= 32


def test_stride_two_is_rejected_and_blocks_build():
    from yolo_somi_b200 import hosting
    with pytest.raises(ValueError, match="stride must be 1"):
        hosting.DCNv3_YOLO(64, 64, 3, 2)
    m = hosting.C3_DCNv3(64, 128, n=2)
    keys = set(m.state_dict())
    assert {"cv1.conv.weight", "cv3.bn.weight", "m.1.cv2.dcn.offset.weight", "m.0.cv1.bn.running_mean"} <= keys
    m = hosting.C2f_DCNv3(64, 100, n=1)          # c = 50: 25 channels per group, constructor must not raise
    assert m.m[0].cv2.dcn.group == 2
    assert isinstance(hosting.C3_DCNv3(32, 32), hosting.C3) and isinstance(hosting.C2f_DCNv3(32, 32), hosting.C2f)


_SNIPPET = '''
from models.common import *
from models.common import ASFF
def parse_model(d, ch):
    for i, (f, n, m, args) in enumerate(d['backbone'] + d['head']):
        m = eval(m) if isinstance(m, str) else m
        if m in [Conv, GhostConv, Bottleneck,
                 C3, C2f]:
            c1, c2 = ch[f], args[0]
            args = [c1, c2, *args[1:]]
            if m in [BottleneckCSP, C3,
                     C2f]:
                args.insert(2, n)
                n = 1
'''


def _lists(text):
    tree = ast.parse(text)
    out = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Compare) and isinstance(node.comparators[0], ast.List):
            out.append([e.id for e in node.comparators[0].elts if isinstance(e, ast.Name)])
    return out


def test_parse_model_patch_on_the_builder_shape():
    import sys
    sys.path.insert(0, str(ROOT / "integration"))
    from apply_parse_model_patch import patch
    out = patch(_SNIPPET)
    assert patch(out) == out                                             # idempotent
    assert "from yolo_somi_b200.hosting import DCNv3_YOLO" in out
    channel, repeat = _lists(out)
    assert {"DCNv3_YOLO", "Bottleneck_DCNv3", "C3_DCNv3", "C2f_DCNv3"} <= set(channel)
    assert {"C3_DCNv3", "C2f_DCNv3"} <= set(repeat) and "DCNv3_YOLO" not in repeat
    ref = Path("/root/reference/models/yolo.py")
    if ref.exists():                                                     # the real builder, in memory only
        text = patch(ref.read_text())
        ast.parse(text)
        lists = _lists(text)
        assert any("C3_DCNv3" in l and "Conv" in l for l in lists) and any("C3_DCNv3" in l and "BottleneckCSP" in l for l in lists)


def test_fuse_for_inference_folds_every_batchnorm_and_keeps_the_outputs():
    """The reference's fuse() (models/yolo.py over utils/torch_utils.py:202-222) restated for the zoo's Conv blocks AND
    the BatchNorm behind a DCNv3 layer (folded into output_proj): same eval-mode function, no BatchNorm left."""
    import copy
    import torch
    from torch import nn
    from yolo_somi_b200 import hosting
    torch.manual_seed(0)
    conv = hosting.Conv(8, 16, 3, 1)
    with torch.no_grad():
        conv.bn.running_mean.normal_(); conv.bn.running_var.uniform_(0.5, 2.0); conv.bn.weight.normal_(); conv.bn.bias.normal_()
    x = torch.randn(2, 8, 10, 10)
    want = conv.eval()(x)
    fused = hosting.fuse_for_inference(copy.deepcopy(conv))
    # the folded bias joins the activation (BiasAct: one pass on the GPU, the plain expression here)
    assert isinstance(fused.bn, nn.Identity) and fused.conv.bias is None and isinstance(fused.act, hosting.BiasAct)
    assert fused.act.kind == "silu" and fused.act.bias.dtype == torch.float32
    assert torch.allclose(fused(x), want, rtol=1e-5, atol=1e-5)
    # an activation BiasAct does not know keeps the bias inside the convolution
    other = hosting.Conv(8, 16, 3, 1, act=nn.ReLU())
    with torch.no_grad():
        other.bn.running_mean.normal_(); other.bn.running_var.uniform_(0.5, 2.0)
    want_other = other.eval()(x)
    fo = hosting.fuse_for_inference(copy.deepcopy(other))
    assert fo.conv.bias is not None and isinstance(fo.act, nn.ReLU) and torch.allclose(fo(x), want_other, rtol=1e-5, atol=1e-5)
    # the wrapper's BatchNorm goes into output_proj: check the affine map itself (the DCNv3 core needs a GPU)
    blk = hosting.DCNv3_YOLO(16, 16, 3)
    with torch.no_grad():
        blk.bn.running_mean.normal_(); blk.bn.running_var.uniform_(0.5, 2.0); blk.bn.weight.normal_(); blk.bn.bias.normal_()
    y = torch.randn(5, 7, 7, 16)                     # what the sampler hands to output_proj (NHWC)
    want = blk.eval().bn(blk.dcn.output_proj(y).permute(0, 3, 1, 2))
    fused = hosting.fuse_for_inference(copy.deepcopy(blk))
    assert isinstance(fused.bn, nn.Identity)
    assert torch.allclose(fused.dcn.output_proj(y).permute(0, 3, 1, 2), want, rtol=1e-5, atol=1e-5)
    assert not any(isinstance(m, nn.BatchNorm2d) for m in fused.modules())
    assert not any(p.requires_grad for p in fused.parameters())
