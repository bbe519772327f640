"""RDSTSR_N -- the RDST variant with the global bottleneck (cat of all RDSTB outputs -> two Linears), reachable from the
same factory with `rdst_global_bottleneck = True`.  Oracle vs reference goldens, wire format and host
logic on CPU; fp32 / bf16 parity of the CUDA path on the GPU."""
import types

import pytest
import torch

import helpers
import rdst_oracle as O
from abi_emulator import emulated_abi


@pytest.mark.parametrize("name", helpers.RDSTN_CASES)
def test_oracle_matches_reference_golden(name):
    c = helpers.load_rdstn_case(name)
    y = O.forward(c["sd"], c["x"], c["scale"])
    assert (y - torch.from_numpy(c["g"]["y"])).abs().max().item() < 2e-5


@pytest.mark.parametrize("name", helpers.RDSTN_CASES)
def test_state_dict_manifest_and_host_logic(name):
    c = helpers.load_rdstn_case(name)
    m = helpers.make_rdstn(c)
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


def test_factory_selects_variant():
    import rdst_b200
    from rdst_b200 import network
    base = dict(patch_size=24, input_channel=1, sr_scale=4.0, swin_patch_size=1, rdst_pre_norm=True,
                rdst_global_bottleneck=True, rdst_global_bottleneck_ratio=1., rdst_global_bottleneck_mode='mlp',
                rdst_feature_last_operation=True, swin_hidden_ratio=2., swin_qkv_bias=True, swin_qk_scale=None,
                swin_drop_rate=0., swin_attn_drop_rate=0., swin_drop_path_rate=0.1, rdst_embed_dim=60,
                rdst_dense_layer_depths=[2] * 8, rdst_num_heads=[6] * 8, rdst_window_size=[8] * 8, rdst_rdb_depths=[3] * 8,
                rdst_layer_norm=True, rdst_ape=False, rdst_patch_norm=True, rdst_use_checkpoint=False,
                rdst_res_connection='1conv', rdst_growth_rate=30, rdst_dense_scale=1., rdst_dim_modify_mode='tail',
                rdst_rdb_residual_scale=1., rdst_global_res_scale=1., rdst_act_in_conv='leaky_relu', rdst_bn_in_conv=None,
                scale_free=False)
    m = rdst_b200.make_RDSTSR(types.SimpleNamespace(**base))
    assert isinstance(m, network.RDSTSR_N)
    assert [k for k, _, _ in helpers.swinir_manifest("rdstn_e1_x4_40x32")] == list(m.state_dict().keys())
    mc = rdst_b200.make_RDSTSR(types.SimpleNamespace(**dict(base, rdst_global_bottleneck_mode='conv')))
    assert tuple(mc.bottleneck[0].weight.shape) == (60, 480, 1, 1) and tuple(mc.bottleneck[1].weight.shape) == (60, 60, 3, 3)
    with pytest.raises(NotImplementedError, match="global_bottleneck_ratio"):
        rdst_b200.make_RDSTSR(types.SimpleNamespace(**dict(base, rdst_global_bottleneck_ratio=0.5)))
    m2 = rdst_b200.make_RDSTSR(types.SimpleNamespace(**dict(base, rdst_global_bottleneck=False)))
    assert type(m2) is network.RDSTSR


@pytest.mark.gpu
@pytest.mark.parametrize("name", helpers.RDSTN_CASES)
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_gpu_matches_reference_golden(name, precision, tol):
    c = helpers.load_rdstn_case(name)
    m = helpers.make_rdstn(c, precision).cuda().eval()
    m.load_state_dict(c["sd"], strict=True)
    with torch.no_grad():
        y = m(c["x"].cuda()).cpu()
    ref = torch.from_numpy(c["g"]["y"])
    assert y.shape == ref.shape and (y - ref).abs().max().item() < tol
    if precision == "bf16":
        target = helpers.realistic_target(ref)
        assert abs(O.psnr(y, target) - O.psnr(ref, target)) < 0.01


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["rdstn_2blk_x2_16x24_b2", "rdstn_conv_3blk_x4_16x16"])
def test_gpu_training_gradients_match_oracle_autograd(name, precision):
    """Gradients through the bottleneck branch; `norm` / `conv_after_body` are unused by this forward and get none."""
    from test_swinir import _grad_check
    c = helpers.load_rdstn_case(name)
    x, s = c["x"], c["scale"]
    target = torch.rand(x.shape[0], 1, x.shape[2] * s, x.shape[3] * s, generator=torch.Generator().manual_seed(5))
    p = {k: (v.clone().double().requires_grad_(True) if v.is_floating_point() else v) for k, v in c["sd"].items()}
    names = [k for k, v in p.items() if v.is_floating_point() and "mean." not in k and "attn_mask" not in k]
    loss_ref = (O.forward(p, x.double(), s) - target.double()).abs().mean()
    g_ref = dict(zip(names, torch.autograd.grad(loss_ref, [p[k] for k in names], allow_unused=True)))
    assert g_ref["norm.weight"] is None and g_ref["conv_after_body.weight"] is None
    m = helpers.make_rdstn(c, precision).cuda().train()
    _grad_check(m, c["sd"], x, target, loss_ref.item(), g_ref, precision == "fp32")
