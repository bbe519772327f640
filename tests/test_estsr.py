"""ESTSR -- the residual-in-residual variant (RRDSTB = RDSTBs + 3x3 conv + shortcut, rdst_variations.py:464-822).
Oracle vs reference goldens, wire format, inference and training host logic on CPU; parity of the CUDA path on the GPU."""
import pytest
import torch

import helpers
import rdst_oracle as O
from abi_emulator import emulated_abi


@pytest.mark.parametrize("name", helpers.ESTSR_CASES)
def test_oracle_matches_reference_golden(name):
    c = helpers.load_estsr_case(name)
    assert (O.forward(c["sd"], c["x"], c["scale"]) - torch.from_numpy(c["g"]["y"])).abs().max().item() < 2e-5


@pytest.mark.parametrize("name", helpers.ESTSR_CASES)
def test_state_dict_manifest_and_host_logic(name):
    c = helpers.load_estsr_case(name)
    m = helpers.make_estsr(c)
    man = helpers.swinir_manifest(name)
    sd = m.state_dict()
    assert [k for k, _, _ in man] == list(sd.keys())
    for k, shape, dt in man:
        assert tuple(sd[k].shape) == tuple(shape) and sd[k].dtype == dt, k
    m.load_state_dict(c["sd"], strict=True)
    with emulated_abi(), torch.no_grad():
        y = m._exec._forward_impl(c["x"])
    ref = torch.from_numpy(c["g"]["y"])
    assert y.shape == ref.shape and (y - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_training_chain_on_cpu(precision):
    from rdst_b200 import autograd
    from test_host_logic_train import _check
    c = helpers.load_estsr_case("estsr_1x3_x2_8x16")
    m = helpers.make_estsr(c, precision)
    target = torch.rand(1, 1, 16, 32, generator=torch.Generator().manual_seed(5))
    _check(m, lambda mod, xx: autograd.forward_with_grad(mod._exec, xx), c["sd"], c["x"], target,
           lambda p, xx: O.forward(p, xx, 2))
    assert m.conv_after_body.weight.grad is None              # registered but unused, as in the reference


@pytest.mark.gpu
@pytest.mark.parametrize("name", helpers.ESTSR_CASES)
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_gpu_matches_reference_golden(name, precision, tol):
    c = helpers.load_estsr_case(name)
    m = helpers.make_estsr(c, precision).cuda().eval()
    m.load_state_dict(c["sd"], strict=True)
    with torch.no_grad():
        y = m(c["x"].cuda()).cpu()
    ref = torch.from_numpy(c["g"]["y"])
    assert y.shape == ref.shape and (y - ref).abs().max().item() < tol * max(1.0, ref.abs().max().item())


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_gpu_training_gradients_match_oracle_autograd(precision):
    from test_swinir import _grad_check
    c = helpers.load_estsr_case("estsr_2x2_x4_16x16_b2")
    x, s = c["x"], c["scale"]
    target = torch.rand(x.shape[0], 1, x.shape[2] * s, x.shape[3] * s, generator=torch.Generator().manual_seed(5))
    p = {k: (v.clone().double().requires_grad_(True) if v.is_floating_point() else v) for k, v in c["sd"].items()}
    names = [k for k, v in p.items() if v.is_floating_point() and "mean." not in k and "attn_mask" not in k]
    loss_ref = (O.forward(p, x.double(), s) - target.double()).abs().mean()
    g_ref = dict(zip(names, torch.autograd.grad(loss_ref, [p[k] for k in names], allow_unused=True)))
    m = helpers.make_estsr(c, precision).cuda().train()
    _grad_check(m, c["sd"], x, target, loss_ref.item(), g_ref, precision == "fp32")
