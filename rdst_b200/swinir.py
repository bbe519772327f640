"""Drop-in replacement for the reference's vanilla SwinIR (lightweight-SR configuration) on the RDST kernels.

SURVEY 8(f) row 2: `networks/swin_transformer_sr.py::SwinIR` / `RSTB` / `swinir_make_model` (reference lines 412-484,
605-869), selected by `feature_generator = 'swinir'` in `models/trans_sr_trainer.py:40` / `trans_sr_tester.py:70`.  It is
the same Swin block arithmetic as RDST at C = 60 (6 heads, window 8, mlp_ratio 2), so the forward is a different launch
sequence over the same librdst_b200 kernels:
    (x - mean) * img_range -> conv_first 3x3 (1->60) -> patch_embed LayerNorm              rdst_head_fwd
    per RSTB: depth x Swin block (fused attention + MLP kernels at C = 60), then
              3x3 conv 60->60 + RSTB residual                                               rdst_conv3x3_fwd[_bf16_tc]
    LayerNorm -> conv_after_body + conv_first skip                                          rdst_layernorm_fwd, conv
    UpsampleOneStep: conv 60 -> s^2 + PixelShuffle(s), / img_range + mean                   rdst_conv3x3_fwd (+ a view)
Same constructor arguments, `forward(x)` and `state_dict` keys as the reference (manifest:
tests/golden/swinir_state_dict_manifest.txt).  Supported envelope: embed_dim 60, 6 heads, window 8, mlp_ratio 2,
in_chans 1, patch_size 1, LayerNorm, no ape, '1conv', upsampler 'pixelshuffledirect', upscale 2/3/4; anything else
raises NotImplementedError.  Training runs through rdst_b200/autograd.py (RSTBFunction chain), fp32 or bf16 GEMMs.
"""
import os

import torch
from torch import nn

from . import _lib, executor, packing
from .network import BasicLayer, PatchEmbed, WINDOW


def _unsupported(what):
    raise NotImplementedError(
        f"rdst_b200.SwinIR: {what} is outside the supported envelope (lightweight SwinIR: embed 60, 6 heads, window 8, "
        "mlp_ratio 2, in_chans 1, '1conv', upsampler 'pixelshuffledirect', upscale 2/3/4).  There is no fallback path.")


class RSTB(nn.Module):
    """Parameter container of swin_transformer_sr.py:412-484 ('1conv')."""

    def __init__(self, dim, input_resolution, depth, num_heads, mlp_ratio, qkv_bias, qk_scale):
        super().__init__()
        self.residual_group = BasicLayer(dim, input_resolution, depth, num_heads, mlp_ratio, qkv_bias, qk_scale)
        self.conv = nn.Conv2d(dim, dim, 3, 1, 1)


class SwinIR(nn.Module):
    """Same constructor signature as the reference SwinIR (swin_transformer_sr.py:632-639)."""

    def __init__(self, img_size=64, patch_size=1, in_chans=3,
                 embed_dim=96, depths=[6, 6, 6, 6], num_heads=[6, 6, 6, 6],
                 window_size=7, mlp_ratio=4., qkv_bias=True, qk_scale=None,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0.1,
                 norm_layer=nn.LayerNorm, ape=False, patch_norm=True,
                 use_checkpoint=False, upscale=2, img_range=1., upsampler='', resi_connection='1conv',
                 precision=None, **kwargs):
        super().__init__()
        if in_chans != 1: _unsupported(f"in_chans={in_chans}")
        if patch_size != 1: _unsupported(f"patch_size={patch_size}")
        if embed_dim != 60: _unsupported(f"embed_dim={embed_dim}")
        if window_size != WINDOW: _unsupported(f"window_size={window_size}")
        if len(depths) != len(num_heads): raise AssertionError("depths / num_heads lengths differ")
        if any(h != 6 for h in num_heads): _unsupported(f"num_heads={num_heads}")
        if float(mlp_ratio) != 2.0: _unsupported(f"mlp_ratio={mlp_ratio}")
        if norm_layer is not nn.LayerNorm: _unsupported("norm_layer other than nn.LayerNorm")
        if ape: _unsupported("absolute position embedding")
        if not patch_norm: _unsupported("patch_norm=False")
        if resi_connection != '1conv': _unsupported(f"resi_connection={resi_connection!r}")
        if upsampler != 'pixelshuffledirect': _unsupported(f"upsampler={upsampler!r}")
        if int(upscale) not in (2, 3, 4): _unsupported(f"upscale={upscale}")
        if drop_rate or attn_drop_rate: _unsupported("dropout > 0")
        if not qkv_bias: _unsupported("qkv_bias=False")

        img = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        self.img_range = img_range
        self.mean = torch.zeros(1, 1, 1, 1)                   # plain attribute, as in the reference (:648)
        self.upscale, self.upsampler = int(upscale), upsampler
        self.num_layers, self.embed_dim, self.num_features = len(depths), embed_dim, embed_dim
        self.ape, self.patch_norm, self.mlp_ratio = ape, patch_norm, mlp_ratio
        self.patches_resolution = img
        self.drop_path_rate = drop_path_rate                  # DropPath is the identity at inference; training raises
        # registration order follows the reference so state_dict() iterates identically
        self.conv_first = nn.Conv2d(in_chans, embed_dim, 3, 1, 1)
        self.patch_embed = PatchEmbed(embed_dim, patch_norm)
        self.layers = nn.ModuleList([RSTB(embed_dim, img, depths[i], num_heads[i], mlp_ratio, qkv_bias, qk_scale)
                                     for i in range(len(depths))])
        self.norm = nn.LayerNorm(embed_dim)
        self.conv_after_body = nn.Conv2d(embed_dim, embed_dim, 3, 1, 1)
        self.upsample = nn.Sequential(nn.Conv2d(embed_dim, self.upscale ** 2 * in_chans, 3, 1, 1),
                                      nn.PixelShuffle(self.upscale))
        self.apply(self._init_weights)
        self.precision = precision or os.environ.get("RDST_B200_PRECISION", "fp32")
        self._exec = SwinIRExecutor(self)

    @staticmethod
    def _init_weights(m):                                     # reference :741-748
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02, a=-2., b=2.)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def set_precision(self, precision):
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision = precision
        return self

    def forward(self, x):
        if not self._exec.bound_to(self):
            self._exec = SwinIRExecutor(self)
        return self._exec.forward(x)

    def extra_repr(self):
        return f"precision={self.precision}, upscale={self.upscale}, backend=librdst_b200 (sm_100a)"


def swinir_make_model(paras):
    """Same contract as the reference factory (swin_transformer_sr.py:830-869), including its img_size rule."""
    upscale = paras.sr_scale
    window_size = paras.sir_window_size
    norm_layer = nn.LayerNorm if paras.sir_layer_norm else nn.Identity
    img_size = int(paras.patch_size // upscale // window_size + 1) * window_size
    return SwinIR(
        img_size=img_size, patch_size=paras.sir_token_size, in_chans=paras.input_channel,
        embed_dim=paras.sir_embed_dim, depths=paras.sir_swintr_layers, num_heads=paras.sir_num_heads,
        window_size=window_size, mlp_ratio=paras.sir_hidden_ratio, qkv_bias=paras.sir_qkv_bias,
        qk_scale=paras.sir_qk_scale, drop_rate=paras.sir_drop_rate, attn_drop_rate=paras.sir_attn_drop_rate,
        drop_path_rate=paras.sir_drop_path_rate, norm_layer=norm_layer, ape=paras.sir_ape,
        patch_norm=paras.sir_patch_norm, use_checkpoint=paras.sir_use_checkpoint, upscale=int(upscale),
        img_range=paras.sir_img_range, upsampler=paras.sir_upsampler, resi_connection=paras.sir_res_connection,
        precision=getattr(paras, "rdst_b200_precision", None))


class SwinIRExecutor(executor.Executor):
    """Launch sequence of the SwinIR forward; reuses the Swin-block / conv launchers and the workspace cache of the RDST
    executor (token-major [T][64] maps, 60 real channels)."""

    def _weights(self, device):
        m = self._module()
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in m.parameters())
        if self._packed is not None and key == self._packed_key:
            return self._packed
        with torch.no_grad():
            f = lambda t: t.detach().float().contiguous()
            id60 = torch.arange(60, device=device)
            P = {"layers": []}
            for layer in m.layers:
                L = {"stl": [packing.pack_stl_tc(packing.pack_stl(b, 60)) for b in layer.residual_group.blocks],
                     "shifts": [b.shift_size for b in layer.residual_group.blocks]}
                L["conv_w"], L["conv_b"] = packing.pack_conv(layer.conv.weight, layer.conv.bias, id60, 64, 64)
                L["conv_img"] = packing.conv_tc_image(L["conv_w"])
                P["layers"].append(L)
            P["head_w"] = f(m.conv_first.weight).reshape(60, 9).contiguous()
            P["head_b"] = f(m.conv_first.bias)
            P["pe_g"], P["pe_b"] = f(m.patch_embed.norm.weight), f(m.patch_embed.norm.bias)
            P["norm_g"], P["norm_b"] = f(m.norm.weight), f(m.norm.bias)
            P["cab_w"], P["cab_b"] = packing.pack_conv(m.conv_after_body.weight, m.conv_after_body.bias, id60, 64, 64)
            P["cab_img"] = packing.conv_tc_image(P["cab_w"])
            up = m.upsample[0]
            P["n_up"] = up.weight.shape[0]
            P["up_w"], P["up_b"] = packing.pack_conv(up.weight, up.bias, id60, 64, 16)
            P["mean"] = float(m.mean.reshape(-1)[0])
        self._packed, self._packed_key = P, key
        return P

    def forward(self, x):
        m = self._module()
        if not x.is_cuda:
            raise RuntimeError("rdst_b200: input must be a CUDA tensor; this package has no CPU path")
        if x.dim() != 4 or x.shape[1] != 1:
            raise ValueError(f"rdst_b200: expected input (B,1,H,W), got {tuple(x.shape)}")
        B, _, H, W = x.shape
        if H % 8 or W % 8:
            raise RuntimeError(f"rdst_b200: H={H}, W={W} must be multiples of the window size 8 "
                               "(the reference fails in window_partition's view for such inputs)")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in m.parameters())):
            from . import autograd
            return autograd.forward_with_grad_swinir(self, x)
        with torch.no_grad(), torch.cuda.device(x.device):
            return self._forward_swinir(x)

    def _forward_swinir(self, x):
        m = self._module()
        dev = x.device
        adt = torch.float32 if m.precision == "fp32" else torch.bfloat16
        dt = _lib.dtype_code(adt)
        B, _, H, W = x.shape
        T = B * H * W
        P = self._weights(dev)
        ws = self._workspace(B, H, W, adt, dev, 1)
        st = _lib.stream_ptr()
        call, ptr = executor.call, executor.ptr
        # [T][64] maps: three trunk buffers rotate (RSTB input / Swin-block ping-pong), carved from the RDST workspace
        v64 = lambda buf: buf.view(-1)[:T * 64].view(T, 64)
        bufs = [v64(ws["D"][0]), v64(ws["D"][1]), v64(ws["Y0"]), v64(ws["Y1"])]
        xin = x.detach().to(torch.float32).contiguous()
        rng = float(m.img_range)
        call("rdst_head_fwd", ptr(xin), rng, -P["mean"] * rng, ptr(P["head_w"]), ptr(P["head_b"]),
             ptr(P["pe_g"]), ptr(P["pe_b"]), ptr(ws["F0"]), 64, ptr(bufs[0]), 64, B, H, W, dt, st)
        cur = 0                                                    # index of the RSTB input
        for L in P["layers"]:
            src = cur
            free = [i for i in range(4) if i != cur]
            for k, (w, shift) in enumerate(zip(L["stl"], L["shifts"])):
                dst = free[k % 2]
                # shift is a constructor-time decision (:188-191); for other input sizes the reference only rebuilds the mask
                # (:254-257), which the kernels evaluate in closed form
                self._stl(bufs[src], 64, bufs[dst], w, shift, B, H, W, ws, dt, st)
                src = dst
            out = free[2] if src != free[2] else free[0]
            self._conv(bufs[src], 64, L["conv_w"], L["conv_img"], L["conv_b"], bufs[cur], 64, bufs[out], 64,
                       B, H, W, 64, 64, 1.0, 0, dt, st)
            cur = out
        call("rdst_layernorm_fwd", ptr(bufs[cur]), 64, ptr(P["norm_g"]), ptr(P["norm_b"]), ptr(ws["FN"]), 64,
             T, 60, 1.0, dt, st)
        self._conv(ws["FN"], 64, P["cab_w"], P["cab_img"], P["cab_b"], ws["F0"], 64, ws["F1"], 64,
                   B, H, W, 64, 64, 1.0, 0, dt, st)
        s = m.upscale
        # the reconstruction conv writes the image: fp32 in both modes (a bf16 store would round the output itself)
        f1 = ws["F1"] if adt == torch.float32 else ws["F1"].float()
        up = torch.empty(T, 16, dtype=torch.float32, device=dev)
        call("rdst_conv3x3_fwd", ptr(f1), 64, ptr(P["up_w"]), ptr(P["up_b"]), None, 0, ptr(up), 16,
             B, H, W, 64, 16, 1.0, 0, _lib.F32, st)
        # PixelShuffle(s) of the s^2 output channels (out channel i*s + j -> sub-pixel (i, j)), / img_range + mean (:795)
        out = up[:, :s * s].reshape(B, H, W, s, s).permute(0, 1, 3, 2, 4).reshape(B, 1, H * s, W * s)
        out = out / rng + P["mean"]
        return out if x.dtype == torch.float32 else out.to(x.dtype)
