"""CPU: the layer-level oracle vs the reference's DCNv3_pytorch outputs (tests/golden/module.npz),
and state_dict key parity of the product module."""
import numpy as np
import pytest
import torch

import cases
from helpers import max_abs, module_golden
from oracle.dcnv3_module_oracle import layer_forward


@pytest.mark.parametrize("mc", cases.MODULE_CASES, ids=lambda m: m.name)
def test_layer_oracle_matches_reference(mc):
    z = module_golden()
    state_np, x_np, grad_np = cases.make_module_state(mc)
    insum = cases.input_checksum([x_np, grad_np] + [state_np[k] for k in sorted(state_np)])
    assert abs(insum - float(z[f"{mc.name}/insum"])) < 1e-6
    state = {k: torch.from_numpy(v).requires_grad_(True) for k, v in state_np.items()}
    x = torch.from_numpy(x_np).requires_grad_(True)
    y = layer_forward(state, x, group=mc.group, kernel_size=mc.kernel_size, stride=mc.stride,
                      pad=mc.pad, dilation=mc.dilation, offset_scale=mc.offset_scale,
                      center_feature_scale=mc.center_feature_scale)
    y.backward(torch.from_numpy(grad_np))
    assert max_abs(y.detach().numpy(), z[f"{mc.name}/y"]) <= 2e-6
    assert max_abs(x.grad.numpy(), z[f"{mc.name}/gx"]) <= 2e-5
    for k, p in state.items():
        want = z[f"{mc.name}/gp/{k}"]
        assert max_abs(p.grad.numpy(), want) <= 2e-5 * max(1.0, float(np.abs(want).max())), k


@pytest.mark.parametrize("mc", cases.MODULE_CASES, ids=lambda m: m.name)
def test_product_module_has_reference_state_dict(mc):
    from yolo_somi_b200.ops_dcnv3.modules import DCNv3
    z = module_golden()
    mod = DCNv3(channels=mc.channels, kernel_size=mc.kernel_size, stride=mc.stride, pad=mc.pad,
                dilation=mc.dilation, group=mc.group, offset_scale=mc.offset_scale,
                center_feature_scale=mc.center_feature_scale)
    assert sorted(mod.state_dict().keys()) == list(z[f"{mc.name}/keys"])
    state_np, _, _ = cases.make_module_state(mc)
    res = mod.load_state_dict({k: torch.from_numpy(v) for k, v in state_np.items()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    # the reference zero-initialises offset/mask and xavier-initialises the projections (:307-315)
    fresh = DCNv3(channels=mc.channels, group=mc.group)
    assert float(fresh.offset.weight.abs().sum() + fresh.mask.bias.abs().sum()) == 0.0
    assert float(fresh.input_proj.bias.abs().sum()) == 0.0 and float(fresh.input_proj.weight.abs().sum()) > 0
