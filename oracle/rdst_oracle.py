"""CPU oracle for the RDST super-resolution hot path.  TEST INFRASTRUCTURE ONLY.

A from-scratch, functional (no nn.Module) restatement of what the reference network computes,
driven directly by a reference-format ``state_dict``.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this file; the product
package ``rdst_b200`` never does (it fails loudly when the CUDA library is missing).

Parity pinning: the reference ships no golden vectors or tests, so this oracle is
pinned against outputs of the reference module itself, generated in the build container by
``oracle/gen_golden.py`` (which imports /root/reference) and committed under ``tests/golden``.
``tests/test_oracle.py`` checks oracle == golden to 2e-5.

Reference lines restated (paths relative to the reference repo):
  networks/rdst_variations.py:1342-1360   RDSTSR.forward            -> forward()
  networks/rdst_variations.py:1326-1340   RDSTSR.forward_features   -> forward()
  networks/rdst_variations.py:438-445     RDSTB.forward             -> rdstb()
  networks/rdst_variations.py:335-341     DenseSTLayer.forward      -> dense_st_layer()
  networks/swin_transformer_sr.py:234-274 SwinTransformerBlock.forward -> swin_block()
  networks/swin_transformer_sr.py:211-232 calculate_mask            -> shift_mask()
  networks/swin_transformer_sr.py:110-141 WindowAttention.forward   -> window_attention()
  networks/swin_transformer_sr.py:32-59   window_partition/reverse  -> to_windows()/from_windows()
  networks/swin_transformer_sr.py:23-29   Mlp.forward               -> inside swin_block()
  networks/common.py:125-148,151-167      UpSampler / MeanShift     -> forward()
"""
import math

import torch
import torch.nn.functional as F

WS = 8          # window size of the supported envelope (E1 ini: rdst_window_size = 8)
HEADS = 6       # rdst_num_heads


def to_windows(x, ws=WS):
    """(B,H,W,C) -> (B*nW, ws*ws, C)   [swin_transformer_sr.py:32-43]"""
    B, H, W, C = x.shape
    if H % ws or W % ws:
        raise ValueError("H and W must be multiples of the window size (reference raises in view())")
    x = x.reshape(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, ws * ws, C)


def from_windows(w, H, W, ws=WS):
    """(B*nW, ws*ws, C) -> (B,H,W,C)   [swin_transformer_sr.py:46-59]"""
    C = w.shape[-1]
    B = w.shape[0] // ((H // ws) * (W // ws))
    x = w.reshape(B, H // ws, W // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(B, H, W, C)


def shift_mask(H, W, ws=WS, shift=WS // 2, dtype=torch.float32):
    """(nW, 64, 64) additive mask of {0,-100} on the SHIFTED frame  [swin_transformer_sr.py:211-232]"""
    region = torch.zeros(H, W, dtype=dtype)
    bounds_h = ((0, H - ws), (H - ws, H - shift), (H - shift, H))
    bounds_w = ((0, W - ws), (W - ws, W - shift), (W - shift, W))
    rid = 0
    for h0, h1 in bounds_h:
        for w0, w1 in bounds_w:
            region[h0:h1, w0:w1] = rid
            rid += 1
    rw = to_windows(region.reshape(1, H, W, 1), ws).reshape(-1, ws * ws)
    diff = rw[:, None, :] - rw[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def rel_pos_index(ws=WS):
    """(64,64) int64 index into the (225, heads) table  [swin_transformer_sr.py:89-98]"""
    ih, iw = torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")
    ih, iw = ih.reshape(-1), iw.reshape(-1)
    dh = ih[:, None] - ih[None, :] + ws - 1
    dw = iw[:, None] - iw[None, :] + ws - 1
    return dh * (2 * ws - 1) + dw


def window_attention(xw, sd, pfx, mask, heads=HEADS):
    """xw: (nWtot, 64, C) normalised tokens  [swin_transformer_sr.py:110-141]"""
    nWt, N, C = xw.shape
    hd = C // heads
    qkv = F.linear(xw, sd[pfx + "qkv.weight"], sd[pfx + "qkv.bias"])
    qkv = qkv.reshape(nWt, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (hd ** -0.5), qkv[1], qkv[2]          # q scaled BEFORE q@k^T (:120)
    att = q @ k.transpose(-2, -1)                             # (nWt, heads, 64, 64)
    idx = sd[pfx + "relative_position_index"].reshape(-1)
    bias = sd[pfx + "relative_position_bias_table"][idx].reshape(N, N, heads).permute(2, 0, 1)
    att = att + bias[None]
    if mask is not None:
        nW = mask.shape[0]
        att = att.reshape(nWt // nW, nW, heads, N, N) + mask[None, :, None].to(att.dtype)
        att = att.reshape(nWt, heads, N, N)
    att = torch.softmax(att, dim=-1)
    out = (att @ v).transpose(1, 2).reshape(nWt, N, C)
    return F.linear(out, sd[pfx + "proj.weight"], sd[pfx + "proj.bias"])


def swin_block(x, H, W, sd, pfx, shift, rnd=None):
    """x: (B, H*W, C)  [swin_transformer_sr.py:234-274]; GELU is the exact erf form (nn.GELU default)."""
    B, L, C = x.shape
    y = F.layer_norm(x, (C,), sd[pfx + "norm1.weight"], sd[pfx + "norm1.bias"], 1e-5).reshape(B, H, W, C)
    if shift:
        y = torch.roll(y, shifts=(-shift, -shift), dims=(1, 2))
        mask = shift_mask(H, W, WS, shift, x.dtype)
    else:
        mask = None
    a = window_attention(to_windows(y), sd, pfx + "attn.", mask)
    a = from_windows(a, H, W)
    if shift:
        a = torch.roll(a, shifts=(shift, shift), dims=(1, 2))
    x = x + a.reshape(B, L, C)
    if rnd is not None:
        x = rnd(x, "stl_mid")
    h = F.layer_norm(x, (C,), sd[pfx + "norm2.weight"], sd[pfx + "norm2.bias"], 1e-5)
    h = F.gelu(F.linear(h, sd[pfx + "mlp.fc1.weight"], sd[pfx + "mlp.fc1.bias"]))
    return x + F.linear(h, sd[pfx + "mlp.fc2.weight"], sd[pfx + "mlp.fc2.bias"])


def dense_st_layer(x, H, W, sd, pfx, dense_scale=1.0, rnd=None):
    """'tail' mode with pre_norm: body(2 STL) -> LN -> Linear(C,growth) -> cat  [rdst_variations.py:335-341];
    'head' mode [:288-295]: LN -> Linear(C,growth) -> body(2 STL at width growth) -> cat."""
    C = x.shape[-1]
    if pfx + "head.0.weight" in sd:
        y = F.layer_norm(x, (C,), sd[pfx + "head.0.weight"], sd[pfx + "head.0.bias"], 1e-5)
        y = F.linear(y, sd[pfx + "head.1.weight"], sd[pfx + "head.1.bias"])
        y = swin_block(y, H, W, sd, pfx + "body.blocks.0.", 0, rnd)
        y = swin_block(y, H, W, sd, pfx + "body.blocks.1.", WS // 2, rnd) * dense_scale
        return torch.cat((x, y), dim=2)
    y = swin_block(x, H, W, sd, pfx + "body.blocks.0.", 0, rnd)
    if rnd is not None:
        y = rnd(y, "stl_out")
    y = swin_block(y, H, W, sd, pfx + "body.blocks.1.", WS // 2, rnd)
    y = F.layer_norm(y, (C,), sd[pfx + "tail.0.weight"], sd[pfx + "tail.0.bias"], 1e-5)
    y = F.linear(y, sd[pfx + "tail.1.weight"], sd[pfx + "tail.1.bias"]) * dense_scale
    if rnd is not None:
        y = rnd(y, "growth")
    return torch.cat((x, y), dim=2)


def conv3x3(x, sd, pfx):
    return F.conv2d(x, sd[pfx + "weight"], sd[pfx + "bias"], padding=1)


def tokens_to_map(x, H, W):
    B, L, C = x.shape
    return x.transpose(1, 2).reshape(B, C, H, W)


def map_to_tokens(x):
    return x.flatten(2).transpose(1, 2)


def rdstb(x, H, W, sd, pfx, n_dstl=3, res_scale=1.0, dense_scale=1.0, rnd=None):
    """3 dense Swin layers -> 3x3 LFF conv (150->60) -> *res_scale -> + shortcut  [rdst_variations.py:438-445]"""
    short = x
    for j in range(n_dstl):
        x = dense_st_layer(x, H, W, sd, f"{pfx}body.{j}.", dense_scale, rnd)
    m = tokens_to_map(x, H, W)
    if pfx + "conv.0.weight" in sd:          # resi_connection = '3conv' (rdst_variations.py:422-427)
        m = F.leaky_relu(conv3x3(m, sd, pfx + "conv.0."), 0.2)
        m = F.leaky_relu(F.conv2d(m, sd[pfx + "conv.2.weight"], sd[pfx + "conv.2.bias"]), 0.2)
        m = conv3x3(m, sd, pfx + "conv.4.")
    else:
        m = conv3x3(m, sd, pfx + "conv.")
    y = map_to_tokens(m) * res_scale
    y = y + short
    if rnd is not None:
        y = rnd(y, "trunk")
    return y


def count_blocks(sd):
    n = 0
    while f"body.{n}.conv.weight" in sd or f"body.{n}.conv.0.weight" in sd:
        n += 1
    return n


def forward(sd, x, sr_scale=4, rnd=None, taps=None, global_res_scale=1.0, feature_last_operation=True,
            rrdb_residual_scale=1.0):
    """Whole network, E1 envelope.  sd: reference state_dict (any float dtype); x: (B,1,H,W).
    `rnd(t, tag)` optionally emulates reduced-precision storage points; `taps` (dict) collects intermediates.
    [rdst_variations.py:1326-1360]"""
    dt = x.dtype
    sd = {k: (v.to(dt) if v.is_floating_point() else v) for k, v in sd.items()}
    B, _, H, W = x.shape
    x = F.conv2d(x, sd["sub_mean.weight"], sd["sub_mean.bias"])
    x0 = conv3x3(x, sd, "head.")
    t = map_to_tokens(x0)
    t = F.layer_norm(t, (t.shape[-1],), sd["patch_embed.norm.weight"], sd["patch_embed.norm.bias"], 1e-5)
    if rnd is not None:
        t = rnd(t, "trunk")
    if taps is not None:
        taps["head"] = x0.clone()
        taps["embed"] = t.clone()
    feats = []
    estsr = "body.0.body.0.conv.weight" in sd
    if estsr:
        # ESTSR [rdst_variations.py:783-812]: body.i = RRDSTB = RDSTBs + 3x3 conv * rrdb_residual_scale + shortcut [:548-555];
        # conv_after_body exists in the state_dict but is not used by that forward
        feature_last_operation = False
        for i in range(count_blocks(sd)):
            short, j = t, 0
            while f"body.{i}.body.{j}.conv.weight" in sd:
                t = rdstb(t, H, W, sd, f"body.{i}.body.{j}.", rnd=rnd)
                j += 1
            t = map_to_tokens(conv3x3(tokens_to_map(t, H, W), sd, f"body.{i}.conv.")) * rrdb_residual_scale + short
    for i in range(0 if estsr else count_blocks(sd)):
        t = rdstb(t, H, W, sd, f"body.{i}.", rnd=rnd)
        feats.append(t)
        if taps is not None:
            taps[f"rdstb{i}"] = t.clone()
    if "bottleneck.0.weight" in sd:
        # RDSTSR_N, global bottleneck 'mlp' [rdst_variations.py:1071-1079,1092-1093]: cat of all RDSTB outputs -> two Linears;
        # `norm` and `conv_after_body` exist in the state_dict but are not used by that forward
        t = torch.cat(feats, 2)
        if sd["bottleneck.0.weight"].dim() == 4:     # 'conv' mode [:1080-1082]: 1x1 conv, then 3x3 conv, on the NCHW map
            fm = tokens_to_map(t, H, W)
            fm = F.conv2d(fm, sd["bottleneck.0.weight"], sd["bottleneck.0.bias"])
            res = conv3x3(fm, sd, "bottleneck.1.") * global_res_scale
        else:
            t = F.linear(t, sd["bottleneck.0.weight"], sd["bottleneck.0.bias"])
            t = F.linear(t, sd["bottleneck.1.weight"], sd["bottleneck.1.bias"])
            res = tokens_to_map(t, H, W) * global_res_scale
    else:
        t = F.layer_norm(t, (t.shape[-1],), sd["norm.weight"], sd["norm.bias"], 1e-5)
        res = tokens_to_map(t, H, W) * global_res_scale
        if feature_last_operation:
            if "conv_after_body.0.weight" in sd:     # '3conv' (rdst_variations.py:1286-1292)
                res = F.leaky_relu(conv3x3(res, sd, "conv_after_body.0."), 0.2)
                res = F.leaky_relu(F.conv2d(res, sd["conv_after_body.2.weight"], sd["conv_after_body.2.bias"]), 0.2)
                res = conv3x3(res, sd, "conv_after_body.4.")
            else:
                res = conv3x3(res, sd, "conv_after_body.")
    res = res + x0
    if rnd is not None:
        res = rnd(res, "feat")
    if taps is not None:
        taps["feat"] = res.clone()
    stage = 0
    s = sr_scale
    while s > 1:
        assert s % 2 == 0, "oracle covers x2/x4 (power-of-two UpSampler branch, common.py:129-132)"
        res = F.pixel_shuffle(conv3x3(res, sd, f"tail.0.{2 * stage}."), 2)
        if rnd is not None:
            res = rnd(res, "up")
        stage += 1
        s //= 2
    out = conv3x3(res, sd, "tail.1." if sr_scale > 1 else "tail.0.")
    return F.conv2d(out, sd["add_mean.weight"], sd["add_mean.bias"])


def psnr(pred, target):
    """10*log10(1/MSE), data_range 1  [metrics/sr_metrics.py:8-9 -> skimage peak_signal_noise_ratio]"""
    mse = torch.mean((pred.double() - target.double()) ** 2).item()
    return float("inf") if mse == 0 else 10.0 * math.log10(1.0 / mse)
