"""oracle/dcnv3_module_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Functional CPU restatement of the reference layer ``DCNv3_pytorch``
(models/ops_dcnv3/modules/dcnv3.py:95-219; forward :183-219) on top of
``oracle.dcnv3_oracle.core_gridsample``.  Takes a plain ``{name: tensor}`` state under the
reference's parameter names.  Pinned against tests/golden/module.npz, which was produced by the
reference's own ``DCNv3_pytorch`` (tests/golden/make_golden.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .dcnv3_oracle import core_gridsample


def layer_forward(state, x, *, group, kernel_size=3, stride=1, pad=1, dilation=1,
                  offset_scale=1.0, center_feature_scale=False, ln_eps=1e-6):
    """x: [N,H,W,C] -> [N,H,W,C]; LN + GELU variant (the reference's defaults)."""
    n, h, w, c = x.shape
    gc = c // group
    proj = F.linear(x, state["input_proj.weight"], state["input_proj.bias"])           # :192
    dwk = state["dw_conv.0.weight"].shape[-1]
    feat = F.conv2d(x.permute(0, 3, 1, 2), state["dw_conv.0.weight"], state["dw_conv.0.bias"],
                    padding=(dwk - 1) // 2, groups=c).permute(0, 2, 3, 1)                # :195-196
    feat = F.gelu(F.layer_norm(feat, (c,), state["dw_conv.1.1.weight"],
                               state["dw_conv.1.1.bias"], ln_eps))
    offset = F.linear(feat, state["offset.weight"], state["offset.bias"])               # :197
    logits = F.linear(feat, state["mask.weight"], state["mask.bias"])                   # :198
    mask = F.softmax(logits.reshape(n, h, w, group, -1), -1).reshape(n, h, w, -1)        # :199
    y = core_gridsample(proj, offset, mask, kernel_size, kernel_size, stride, stride, pad, pad,
                        dilation, dilation, group, gc, offset_scale)                    # :201-208
    if center_feature_scale:                                                            # :209-215
        cfs = F.linear(feat, state["center_feature_scale_proj_weight"],
                       state["center_feature_scale_proj_bias"]).sigmoid()
        cfs = cfs[..., None].repeat(1, 1, 1, 1, gc).flatten(-2)
        y = y * (1 - cfs) + proj * cfs
    return F.linear(y, state["output_proj.weight"], state["output_proj.bias"])          # :216
