"""Shared test helpers: golden loading, reference-format state_dict skeletons, module construction."""
import os

import numpy as np
import torch

import rdst_oracle as O
from synth_weights import fill_state_dict, synth_input

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ["e1_x4_64x64", "e1_x4_16x24_b2", "e_x4_40x32", "e1_x2_24x24", "e2blk_x4_8x8"]


def manifest():
    rows = []
    with open(os.path.join(GOLDEN, "e1_state_dict_manifest.txt")) as f:
        for line in f:
            k, shape, dt = line.rstrip("\n").split("\t")
            rows.append((k, eval(shape), getattr(torch, dt)))
    return rows


def skeleton_state_dict(blocks=8, scale=4):
    """Reference-format state_dict (zeros) for a `blocks`-RDSTB model, built from the committed manifest."""
    sd = {}
    for k, shape, dt in manifest():
        if k.startswith("body.") and int(k.split(".")[1]) >= blocks:
            continue
        if scale == 2 and k.startswith("tail.0.2."):
            continue
        sd[k] = torch.zeros(shape, dtype=dt)
    sd["sub_mean.weight"][:] = 1
    sd["add_mean.weight"][:] = 1
    for k in sd:
        if k.endswith("relative_position_index"):
            sd[k] = O.rel_pos_index()
        if k.endswith("attn_mask"):
            sd[k] = O.shift_mask(24, 24)
    return sd


def load_case(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    blocks, scale = int(g["meta_blocks"]), int(g["meta_scale"])
    sd = fill_state_dict(skeleton_state_dict(blocks, scale), int(g["meta_wseed"]), bool(g["meta_perturbed"]))
    x = synth_input(tuple(int(v) for v in g["shape"]), int(g["meta_xseed"]))
    return dict(g=g, blocks=blocks, scale=scale, sd=sd, x=x)


def realistic_target(ref, seed=123, sigma=0.0224):
    """A ground-truth stand-in at realistic SR quality: the reference output + Gaussian noise at ~33 dB PSNR.  Against
    such a target a 0.01 dB PSNR tolerance bounds the RMS error of the tested output near 1e-3 of the data range (against
    a uniform-random target the check could never fail).  PSNR is defined relative to the data range (1 for the [0,1]
    images of the north_star bar); the synthetic-weight SwinIR / RDSTSR_N fixtures span about +-1.2, so the noise level is
    scaled by the reference's own range when that exceeds 1 -- the same rule the max-abs bar of those tests uses."""
    rng = max(1.0, float(ref.max() - ref.min()))
    return ref + sigma * rng * torch.randn(ref.shape, generator=torch.Generator().manual_seed(seed))


def make_module(blocks=8, scale=4, precision="fp32", img_size=24):
    import rdst_b200
    return rdst_b200.RDSTSR(img_size=img_size, sr_scale=scale, dense_layer_depths=[2] * blocks,
                            num_heads=[6] * blocks, window_size=[8] * blocks, rdb_depths=[3] * blocks,
                            mlp_ratio=2., pre_norm=True, feature_last_operation=True, precision=precision)


# ---- vanilla SwinIR: fixtures from oracle/gen_golden_swinir.py ----
SWINIR_CASES = ["swinir_ini_x4_40x32", "swinir_x2_16x24_b2", "swinir_x3_8x8"]


def swinir_manifest(name):
    rows = []
    with open(os.path.join(GOLDEN, name + "_manifest.txt")) as f:
        for line in f:
            k, shape, dt = line.rstrip("\n").split("\t")
            rows.append((k, eval(shape), getattr(torch, dt)))
    return rows


def load_swinir_case(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    img = int(g["meta_img_size"])
    sd = {}
    for k, shape, dt in swinir_manifest(name):
        sd[k] = torch.zeros(shape, dtype=dt)
        if k.endswith("relative_position_index"):
            sd[k] = O.rel_pos_index()
        if k.endswith("attn_mask"):
            sd[k] = O.shift_mask(img, img)
    sd = fill_state_dict(sd, int(g["meta_wseed"]), True)
    x = synth_input(tuple(int(v) for v in g["shape"]), int(g["meta_xseed"]))
    return dict(g=g, sd=sd, x=x, img_size=img, upscale=int(g["meta_upscale"]), depths=[int(d) for d in g["depths"]])


def make_swinir(c, precision="fp32"):
    from rdst_b200 import swinir
    return swinir.SwinIR(img_size=c["img_size"], patch_size=1, in_chans=1, embed_dim=60, depths=c["depths"],
                         num_heads=[6] * len(c["depths"]), window_size=8, mlp_ratio=2., upscale=c["upscale"], img_range=1.,
                         upsampler="pixelshuffledirect", resi_connection="1conv", precision=precision,
                         drop_path_rate=c.get("drop_path_rate", 0.0))


# ---- RDSTSR_N (global bottleneck): fixtures from oracle/gen_golden_rdstn.py ----
RDSTN_CASES = ["rdstn_e1_x4_40x32", "rdstn_2blk_x2_16x24_b2", "rdstn_conv_3blk_x4_16x16"]


def load_rdstn_case(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    blocks, scale = int(g["meta_blocks"]), int(g["meta_scale"])
    sd = {}
    for k, shape, dt in swinir_manifest(name):
        sd[k] = torch.zeros(shape, dtype=dt)
        if k.endswith("relative_position_index"):
            sd[k] = O.rel_pos_index()
        if k.endswith("attn_mask"):
            sd[k] = O.shift_mask(24, 24)
    sd["sub_mean.weight"][:] = 1
    sd["add_mean.weight"][:] = 1
    sd = fill_state_dict(sd, int(g["meta_wseed"]), True)
    x = synth_input(tuple(int(v) for v in g["shape"]), int(g["meta_xseed"]))
    mode = None if "bottleneck.0.weight" not in sd else ("conv" if sd["bottleneck.0.weight"].dim() == 4 else "mlp")
    return dict(g=g, sd=sd, x=x, blocks=blocks, scale=scale, mode=mode)


def make_rdstn(c, precision="fp32"):
    from rdst_b200 import network
    b = c["blocks"]
    return network.RDSTSR_N(img_size=24, sr_scale=c["scale"], dense_layer_depths=[2] * b, num_heads=[6] * b,
                            window_size=[8] * b, rdb_depths=[3] * b, mlp_ratio=2., pre_norm=True,
                            global_bottleneck_mode=c.get("mode", "mlp"), precision=precision)


# ---- ESTSR (residual-in-residual RDSTBs): fixtures from oracle/gen_golden_estsr.py ----
ESTSR_CASES = ["estsr_2x2_x4_16x16_b2", "estsr_1x3_x2_8x16"]


def load_estsr_case(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    sd = {}
    for k, shape, dt in swinir_manifest(name):
        sd[k] = torch.zeros(shape, dtype=dt)
        if k.endswith("relative_position_index"):
            sd[k] = O.rel_pos_index()
        if k.endswith("attn_mask"):
            sd[k] = O.shift_mask(24, 24)
    sd["sub_mean.weight"][:] = 1
    sd["add_mean.weight"][:] = 1
    sd = fill_state_dict(sd, int(g["meta_wseed"]), True)
    x = synth_input(tuple(int(v) for v in g["shape"]), int(g["meta_xseed"]))
    return dict(g=g, sd=sd, x=x, n_rr=int(g["meta_n_rr"]), n_rd=int(g["meta_n_rd"]), scale=int(g["meta_scale"]))


def make_estsr(c, precision="fp32"):
    import rdst_b200
    n = c["n_rr"]
    return rdst_b200.ESTSR(img_size=24, sr_scale=c["scale"], dense_layer_depths=[2] * n, num_heads=[6] * n, window_size=[8] * n,
                           rdb_depths=[3] * n, rrdb_depths=[c["n_rd"]] * n, num_rrdb_blocks=n, mlp_ratio=2., pre_norm=True,
                           precision=precision)


# ---- RDSTSR with resi_connection = '3conv' (SURVEY 8f row 3): fixtures from oracle/gen_golden_3conv.py ----
C3_CASES = ["rdst3conv_2blk_x4_16x24_b2", "rdst3conv_3blk_x2_24x24"]


def load_3conv_case(name):
    c = load_rdstn_case(name)           # same fixture format (manifest + npz)
    c.pop("mode", None)
    return c


def make_3conv(c, precision="fp32"):
    import rdst_b200
    b = c["blocks"]
    return rdst_b200.RDSTSR(img_size=24, sr_scale=c["scale"], dense_layer_depths=[2] * b, num_heads=[6] * b, window_size=[8] * b,
                            rdb_depths=[3] * b, mlp_ratio=2., pre_norm=True, feature_last_operation=True,
                            resi_connection="3conv", precision=precision)


# ---- RDSTSR with dim_modify_mode = 'head' (SURVEY 8f row 3): fixtures from oracle/gen_golden_headmode.py ----
HEADMODE_CASES = ["rdsthead_2blk_x4_16x24_b2", "rdsthead_3blk_x2_24x24"]


def make_headmode(c, precision="fp32"):
    import rdst_b200
    b = c["blocks"]
    return rdst_b200.RDSTSR(img_size=24, sr_scale=c["scale"], dense_layer_depths=[2] * b, num_heads=[6] * b, window_size=[8] * b,
                            rdb_depths=[3] * b, mlp_ratio=2., pre_norm=True, feature_last_operation=True,
                            dim_modify_mode="head", precision=precision)
