"""CPU tests of the TRAINING host logic: the chain of autograd Functions (head / RDSTB / RSTB / bottleneck / tail), the
packed-weight layouts and pack descriptors, buffer strides and the flow of gradients back to the reference-named parameters
-- with every C-ABI entry point replaced by its contract restatement (tests/abi_emulator.py) -- against torch.autograd
through the fp64 oracle.  The CUDA kernels themselves are checked against the same contracts by the `-m gpu` tests."""
import pytest
import torch

import helpers
import rdst_oracle as O
import swinir_oracle as SO
from abi_emulator import emulated_abi
from synth_weights import fill_state_dict


def _check(m, fwd, sd, x, target, oracle, skip=("mean.", "attn_mask")):
    p = {k: (v.clone().double().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    names = [k for k, v in p.items() if v.is_floating_point() and not any(s in k for s in skip)]
    loss_ref = (oracle(p, x.double()) - target.double()).abs().mean()
    g_ref = dict(zip(names, torch.autograd.grad(loss_ref, [p[k] for k in names], allow_unused=True)))
    m.load_state_dict(sd, strict=True)
    m.train()
    calls = []
    with emulated_abi(calls):
        out = fwd(m, x)
        assert out.requires_grad
        loss = torch.nn.functional.l1_loss(out, target)
        loss.backward()
    assert abs(loss.item() - loss_ref.item()) < 1e-5
    params = dict(m.named_parameters())
    for k, gr in g_ref.items():
        if gr is None:
            assert params[k].grad is None, k
            continue
        assert params[k].grad is not None, k
        rel = ((params[k].grad.double() - gr).norm() / (gr.norm() + 1e-12)).item()
        assert rel < 1e-4, (k, rel)
    return calls


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_rdst_training_chain_on_cpu(precision):
    """precision='bf16' takes the tensor-core entry points (padded q|k|v rows, fused GELU / LayerNorm operands, saved
    log-sum-exp, LayerNorm backward in the GEMM epilogue); their stand-ins do not round, so the same tolerance applies."""
    from rdst_b200 import autograd
    sd = fill_state_dict(helpers.skeleton_state_dict(1, 4), 11, True)
    m = helpers.make_module(1, 4, precision)
    x = torch.rand(2, 1, 16, 24, generator=torch.Generator().manual_seed(4))
    target = torch.rand(2, 1, 64, 96, generator=torch.Generator().manual_seed(5))
    calls = _check(m, lambda mod, xx: autograd.forward_with_grad(mod._exec, xx), sd, x, target,
                   lambda p, xx: O.forward(p, xx, 4))
    assert calls.count("rdst_pack_linear_batch") == 2            # one forward + one backward launch per RDSTB
    if precision == "fp32":
        assert calls.count("rdst_window_attention_bwd") == 6 and "rdst_gelu_fwd" in calls and "rdst_gemm_tc" not in calls
    else:
        assert calls.count("rdst_window_attention_tc_bwd") == 6 and calls.count("rdst_gemm_tc_lnbwd") == 15
        assert not {"rdst_linear_fwd", "rdst_gemm_tn_acc", "rdst_gelu_fwd", "rdst_gelu_bwd", "rdst_lnhat_fwd", "rdst_lnhat_bwd",
                    "rdst_conv3x3_fwd", "rdst_window_attention_fwd"} & set(calls)      # nothing falls back to CUDA cores
    assert m.sub_mean.weight.grad is None and m.add_mean.bias.grad is None


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_swinir_training_chain_on_cpu(precision):
    from rdst_b200 import autograd
    c = helpers.load_swinir_case("swinir_x2_16x24_b2")
    m = helpers.make_swinir(c, precision)
    target = torch.rand(2, 1, 32, 48, generator=torch.Generator().manual_seed(5))
    _check(m, lambda mod, xx: autograd.forward_with_grad_swinir(mod._exec, xx), c["sd"], c["x"], target,
           lambda p, xx: SO.forward(p, xx, 2), skip=("attn_mask",))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["rdstn_2blk_x2_16x24_b2", "rdstn_conv_3blk_x4_16x16"])
def test_rdstn_training_chain_on_cpu(name, precision):
    from rdst_b200 import autograd
    c = helpers.load_rdstn_case(name)
    m = helpers.make_rdstn(c, precision)
    s, x = c["scale"], c["x"]
    target = torch.rand(x.shape[0], 1, x.shape[2] * s, x.shape[3] * s, generator=torch.Generator().manual_seed(5))
    _check(m, lambda mod, xx: autograd.forward_with_grad(mod._exec, xx), c["sd"], x, target,
           lambda p, xx: O.forward(p, xx, s))
    assert m.norm.weight.grad is None and m.conv_after_body.weight.grad is None      # unused by this forward


def test_training_guards_fail_loudly():
    """parameters that are not contiguous fp32 and SwinIR stochastic depth must raise, not train differently."""
    from rdst_b200 import autograd
    m = helpers.make_module(1, 4, "bf16").train()
    m.body[0].conv.weight.data = m.body[0].conv.weight.data.to(torch.bfloat16)
    with pytest.raises(TypeError, match="contiguous float32"):
        autograd.forward_with_grad(m._exec, torch.rand(1, 1, 8, 8))
    c = helpers.load_swinir_case("swinir_x2_16x24_b2")
    ms = helpers.make_swinir(dict(c, drop_path_rate=0.1), "fp32").train()
    with pytest.raises(NotImplementedError, match="stochastic depth"):
        autograd.forward_with_grad_swinir(ms._exec, c["x"])
    ms.eval()
    assert ms.drop_path_rate == 0.1          # stored for interface parity; identity at inference, as in the reference


def test_mlp_ratio_outside_envelope_raises():
    import rdst_b200
    ok = dict(img_size=24, sr_scale=4, dense_layer_depths=[2] * 2, num_heads=[6] * 2, window_size=[8] * 2,
              rdb_depths=[3] * 2, pre_norm=True, feature_last_operation=True)
    with pytest.raises(NotImplementedError, match="mlp_ratio"):
        rdst_b200.RDSTSR(mlp_ratio=4., **ok)
    with pytest.raises(NotImplementedError, match="mlp_ratio"):
        rdst_b200.network.RDSTSR_N(mlp_ratio=4., global_bottleneck=True, **{k: v for k, v in ok.items() if k != "feature_last_operation"})
