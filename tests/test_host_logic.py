"""CPU tests of the host side: drop-in surface, weight packing, launch order -- checked against the golden
vectors produced by the reference module, with the kernels replaced by their contract restatements."""
import copy

import numpy as np
import pytest
import torch

import helpers
from abi_emulator import emulated_abi


def _run_emulated(mod, x):
    with emulated_abi(), torch.no_grad():
        return mod._exec._forward_impl(x)


@pytest.mark.parametrize("name", helpers.CASES)
def test_forward_matches_reference_golden(name):
    c = helpers.load_case(name)
    m = helpers.make_module(c["blocks"], c["scale"], "fp32")
    m.load_state_dict(c["sd"], strict=True)
    y = _run_emulated(m, c["x"])
    ref = torch.from_numpy(c["g"]["y"])
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() < 2e-5


def test_state_dict_manifest_exact():
    m = helpers.make_module(8, 4)
    sd = m.state_dict()
    man = helpers.manifest()
    assert [k for k, _, _ in man] == list(sd.keys())
    for k, shape, dt in man:
        assert tuple(sd[k].shape) == tuple(shape) and sd[k].dtype == dt, k
    assert sum(p.numel() for p in m.parameters()) == 4464965
    assert sum(p.numel() for p in m.parameters() if p.requires_grad) == 4464961
    assert not m.sub_mean.weight.requires_grad and not m.add_mean.bias.requires_grad


def test_repack_after_weight_update():
    c = helpers.load_case("e2blk_x4_8x8")
    m = helpers.make_module(c["blocks"], c["scale"])
    m.load_state_dict(c["sd"])
    y0 = _run_emulated(m, c["x"]).clone()
    with torch.no_grad():
        m.body[0].body[1].body.blocks[1].mlp.fc2.weight.mul_(2.0)      # (a uniform bias shift would be removed by the LN)
    y1 = _run_emulated(m, c["x"])
    assert (y1 - y0).abs().max().item() > 1e-4          # cache must notice the in-place update
    m.load_state_dict(c["sd"])
    y2 = _run_emulated(m, c["x"])
    assert torch.equal(y2, y0)


def test_deepcopy_rebinds_executor():
    m = helpers.make_module(2, 4)
    m2 = copy.deepcopy(m)
    assert not m2._exec.bound_to(m2)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        m2(torch.zeros(1, 1, 8, 8))
    assert m2._exec.bound_to(m2)


def test_errors_mirror_reference():
    m = helpers.make_module(2, 4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 1, 8, 8))
    import rdst_b200
    with pytest.raises(NotImplementedError):
        rdst_b200.RDSTSR(img_size=24, window_size=[4] * 4)
    ok = dict(img_size=24, sr_scale=4, dense_layer_depths=[2] * 2, num_heads=[6] * 2, window_size=[8] * 2,
              rdb_depths=[3] * 2, mlp_ratio=2., pre_norm=True, feature_last_operation=True)
    with pytest.raises(ValueError):
        rdst_b200.RDSTSR(act_in_conv="tanh", **ok)
    with pytest.raises(ValueError):
        rdst_b200.RDSTSR(mean=[0., 0.], std=[1.], **ok)


def test_make_rdstsr_reads_reference_paras():
    class P:  # the attribute names of config_files/RDST_E1_OASIS_example_SRx4.ini
        patch_size = 24; input_channel = 1; sr_scale = 4.0; swin_patch_size = 1
        rdst_pre_norm = True; rdst_global_bottleneck = False; rdst_global_bottleneck_ratio = 1.
        rdst_feature_last_operation = True; swin_hidden_ratio = 2.; swin_qkv_bias = True; swin_qk_scale = None
        swin_drop_rate = 0.; swin_attn_drop_rate = 0.; swin_drop_path_rate = 0.1; rdst_embed_dim = 60
        rdst_dense_layer_depths = [2] * 8; rdst_num_heads = [6] * 8; rdst_window_size = [8] * 8
        rdst_rdb_depths = [3] * 8; rdst_layer_norm = True; rdst_ape = False; rdst_patch_norm = True
        rdst_use_checkpoint = False; rdst_res_connection = '1conv'; rdst_growth_rate = 30; rdst_dense_scale = 1.
        rdst_dim_modify_mode = 'tail'; rdst_rdb_residual_scale = 1.; rdst_global_res_scale = 1.
        rdst_act_in_conv = 'leaky_relu'; rdst_bn_in_conv = None; scale_free = False
    import rdst_b200
    m = rdst_b200.make_RDSTSR(P())
    assert len(m.state_dict()) == 826 and m.sr_scale == 4
    m = rdst_b200.make_RDSTSR(P(), mean=[0.5], std=[2.0])
    assert float(m.sub_mean.bias) == -0.25 and float(m.add_mean.weight) == 2.0
