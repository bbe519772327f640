"""CPU oracle for the vanilla SwinIR path.  TEST INFRASTRUCTURE ONLY -- same rules as rdst_oracle.py.

Functional restatement of networks/swin_transformer_sr.py::SwinIR.forward / forward_features (:768-812, lightweight
'pixelshuffledirect' branch :795-799), RSTB.forward (:471-472) and UpsampleOneStep (:583-596), driven by a
reference-format state_dict; the Swin block itself is rdst_oracle.swin_block.  Pinned against outputs of the reference
module (oracle/gen_golden_swinir.py -> tests/golden/swinir_*.npz, tests/test_oracle.py)."""
import torch
import torch.nn.functional as F

import rdst_oracle as O


def forward(sd, x, upscale, img_range=1.0, mean=0.0, rnd=None):
    dt = x.dtype
    sd = {k: (v.to(dt) if v.is_floating_point() else v) for k, v in sd.items()}
    B, _, H, W = x.shape
    x = (x - mean) * img_range
    x0 = O.conv3x3(x, sd, "conv_first.")
    t = O.map_to_tokens(x0)
    t = F.layer_norm(t, (t.shape[-1],), sd["patch_embed.norm.weight"], sd["patch_embed.norm.bias"], 1e-5)
    i = 0
    while f"layers.{i}.conv.weight" in sd:
        short, j = t, 0
        while f"layers.{i}.residual_group.blocks.{j}.norm1.weight" in sd:
            pfx = f"layers.{i}.residual_group.blocks.{j}."
            # the shift of a block is decided in its constructor (:188-191) and is visible as the attn_mask buffer
            t = O.swin_block(t, H, W, sd, pfx, O.WS // 2 if pfx + "attn_mask" in sd else 0, rnd)
            j += 1
        t = O.map_to_tokens(O.conv3x3(O.tokens_to_map(t, H, W), sd, f"layers.{i}.conv.")) + short
        i += 1
    t = F.layer_norm(t, (t.shape[-1],), sd["norm.weight"], sd["norm.bias"], 1e-5)
    res = O.conv3x3(O.tokens_to_map(t, H, W), sd, "conv_after_body.") + x0
    out = F.pixel_shuffle(O.conv3x3(res, sd, "upsample.0."), upscale)
    return out / img_range + mean
