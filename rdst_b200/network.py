"""Drop-in replacement for the reference RDST super-resolution network.

Mirrors the constructor arguments, ``forward(x, sr_scale=None)`` signature and ``state_dict`` key set of
``networks/rdst_variations.py::RDSTSR`` / ``make_RDSTSR`` (reference lines 1115-1457) so that
``models/trans_sr_trainer.py`` / ``trans_sr_tester.py`` load it unchanged.  The nn.Module tree below exists only
to own parameters under the reference's names; all arithmetic is done by hand-written CUDA kernels in
librdst_b200.so (see rdst_b200/executor.py).  There is no CPU path: a CPU tensor raises.
"""
import os

import torch
from torch import nn

from . import executor

WINDOW = 8


class WindowAttention(nn.Module):
    """Parameter container for swin_transformer_sr.py:62-141 (qkv, proj, relative-position table/index)."""

    def __init__(self, dim, num_heads, qkv_bias=True, qk_scale=None):
        super().__init__()
        self.dim, self.num_heads = dim, num_heads
        self.window_size = (WINDOW, WINDOW)
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * WINDOW - 1) ** 2, num_heads))
        r = torch.arange(WINDOW)
        ih, iw = (t.reshape(-1) for t in torch.meshgrid(r, r, indexing="ij"))
        index = (ih[:, None] - ih[None, :] + WINDOW - 1) * (2 * WINDOW - 1) + (iw[:, None] - iw[None, :] + WINDOW - 1)
        self.register_buffer("relative_position_index", index)
        self.qkv = nn.Linear(dim, 3 * dim, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02, a=-2., b=2.)

    def extra_repr(self):
        return f"dim={self.dim}, window_size={self.window_size}, num_heads={self.num_heads}"


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)


def shift_mask(h, w, shift):
    """(nW,64,64) buffer of {0,-100}: kept only for state_dict parity (swin_transformer_sr.py:211-232);
    the kernels evaluate the same mask in closed form from window coordinates."""
    region = torch.zeros(h, w)
    rid = 0
    for h0, h1 in ((0, h - WINDOW), (h - WINDOW, h - shift), (h - shift, h)):
        for w0, w1 in ((0, w - WINDOW), (w - WINDOW, w - shift), (w - shift, w)):
            region[h0:h1, w0:w1] = rid
            rid += 1
    rw = region.reshape(h // WINDOW, WINDOW, w // WINDOW, WINDOW).permute(0, 2, 1, 3).reshape(-1, WINDOW * WINDOW)
    diff = rw[:, None, :] - rw[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


class SwinTransformerBlock(nn.Module):
    def __init__(self, dim, input_resolution, num_heads, shift_size, mlp_ratio, qkv_bias, qk_scale):
        super().__init__()
        self.dim, self.input_resolution, self.num_heads = dim, tuple(input_resolution), num_heads
        self.window_size, self.shift_size, self.mlp_ratio = WINDOW, shift_size, mlp_ratio
        if min(self.input_resolution) <= WINDOW:          # reference :188-191
            self.shift_size = 0
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, num_heads, qkv_bias, qk_scale)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        mask = shift_mask(*self.input_resolution, self.shift_size) if self.shift_size > 0 else None
        self.register_buffer("attn_mask", mask)

    def extra_repr(self):
        return (f"dim={self.dim}, input_resolution={self.input_resolution}, num_heads={self.num_heads}, "
                f"window_size={self.window_size}, shift_size={self.shift_size}, mlp_ratio={self.mlp_ratio}")


class BasicLayer(nn.Module):
    def __init__(self, dim, input_resolution, depth, num_heads, mlp_ratio, qkv_bias, qk_scale):
        super().__init__()
        self.blocks = nn.ModuleList([
            SwinTransformerBlock(dim, input_resolution, num_heads, 0 if i % 2 == 0 else WINDOW // 2,
                                 mlp_ratio, qkv_bias, qk_scale) for i in range(depth)])


class DenseSTLayer(nn.Module):
    """rdst_variations.py:246-341 with pre_norm.  'tail' mode: body (2 Swin blocks at the input width) -> LN -> Linear(C, growth).
    'head' mode (:288-304): LN -> Linear(C, growth) -> body (2 Swin blocks at width `growth`)."""

    def __init__(self, input_dim, input_resolution, depth, num_heads, mlp_ratio, qkv_bias, qk_scale, growth_rate,
                 dim_modify_mode='tail'):
        super().__init__()
        self.input_dim, self.growth_rate, self.dim_modify_mode = input_dim, growth_rate, dim_modify_mode
        if dim_modify_mode == 'head':
            self.head = nn.Sequential(nn.LayerNorm(input_dim), nn.Linear(input_dim, growth_rate))
            self.body = BasicLayer(growth_rate, input_resolution, depth, num_heads, mlp_ratio, qkv_bias, qk_scale)
        else:
            self.tail = nn.Sequential(nn.LayerNorm(input_dim), nn.Linear(input_dim, growth_rate))
            self.body = BasicLayer(input_dim, input_resolution, depth, num_heads, mlp_ratio, qkv_bias, qk_scale)


class RDSTB(nn.Module):
    """rdst_variations.py:354-445: dense Swin layers + 3x3 local-feature-fusion conv + residual."""

    def __init__(self, input_dim, input_resolution, layer_depth, num_heads, mlp_ratio, qkv_bias, qk_scale,
                 growth_rate, num_blocks, resi_connection='1conv', dim_modify_mode='tail'):
        super().__init__()
        self.body = nn.ModuleList()
        dim = input_dim
        for _ in range(num_blocks):
            self.body.append(DenseSTLayer(dim, input_resolution, layer_depth, num_heads, mlp_ratio,
                                          qkv_bias, qk_scale, growth_rate, dim_modify_mode))
            dim += growth_rate
        if resi_connection == '1conv':
            self.conv = nn.Conv2d(dim, input_dim, 3, 1, 1)
        else:                                       # '3conv' (rdst_variations.py:422-427): bottleneck fusion, dim // 4 channels
            self.conv = nn.Sequential(nn.Conv2d(dim, dim // 4, 3, 1, 1), nn.LeakyReLU(negative_slope=0.2, inplace=True),
                                      nn.Conv2d(dim // 4, dim // 4, 1, 1, 0), nn.LeakyReLU(negative_slope=0.2, inplace=True),
                                      nn.Conv2d(dim // 4, input_dim, 3, 1, 1))
        self.resi_connection = resi_connection


class PatchEmbed(nn.Module):
    def __init__(self, embed_dim, norm):
        super().__init__()
        self.norm = nn.LayerNorm(embed_dim) if norm else None


class MeanShift(nn.Conv2d):
    """Frozen 1x1 conv (networks/common.py:151-167); folded into the head / last conv kernels as a scalar affine."""

    def __init__(self, mean, std, mode):
        nc = len(mean)
        super().__init__(nc, nc, kernel_size=1)
        std_t = torch.tensor(std, dtype=torch.float32)
        mean_t = torch.tensor(mean, dtype=torch.float32)
        if mode == "sub":
            self.weight.data = torch.eye(nc).view(nc, nc, 1, 1) / std_t.view(nc, 1, 1, 1)
            self.bias.data = -mean_t / std_t
        else:
            self.weight.data = torch.eye(nc).view(nc, nc, 1, 1) * std_t.view(nc, 1, 1, 1)
            self.bias.data = mean_t.clone()
        for p in self.parameters():
            p.requires_grad = False


def _unsupported(what):
    raise NotImplementedError(
        f"rdst_b200: {what} is outside the supported envelope (RDST-E1 family: window 8, 6 heads, embed 60, "
        "growth 30, 3 dense layers of depth 2, 'tail'+pre_norm, '1conv', LayerNorm, no ape, scale 2/4). "
        "There is deliberately no fallback path.")


class RDSTSR(nn.Module):
    """Same constructor signature as the reference RDSTSR (rdst_variations.py:1142-1157)."""

    def __init__(self, img_size=48, patch_size=1, in_chans=1, sr_scale=2, embed_dim=60,
                 dense_layer_depths=[2, 2, 2, 2], num_heads=[6, 6, 6, 6],
                 window_size=[4, 4, 4, 4], rdb_depths=[3, 3, 3, 3],
                 mlp_ratio=4., qkv_bias=True, qk_scale=None,
                 drop_rate=0., attn_drop=0., drop_path_rate=0.,
                 norm_layer=nn.LayerNorm, ape=False, patch_norm=True,
                 use_checkpoint=False, resi_connection='1conv',
                 growth_rate=30, dense_scale=1., dim_modify_mode='tail',
                 rdb_residual_scale=1., global_res_scale=1.,
                 mean=None, std=None,
                 act_in_conv='leaky_relu', bn_in_conv=None,
                 scale_free=False, scale_embedding=False,
                 pre_norm=False, feature_last_operation=False,
                 precision=None):
        super().__init__()
        n = len(rdb_depths)
        if not (len(window_size) == len(num_heads) == len(dense_layer_depths) == n):
            raise AssertionError("rdb_depths / window_size / num_heads / dense_layer_depths lengths differ")
        if act_in_conv not in ('relu', 'leaky_relu', 'prelu'):
            raise ValueError('Invalid activation {}, should be one of [relu, leaky_relu, prelu]'.format(act_in_conv))
        # ---- envelope checks: anything else must fail loudly, not diverge silently ----
        if in_chans != 1: _unsupported(f"in_chans={in_chans}")
        if patch_size != 1: _unsupported(f"patch_size={patch_size}")
        if embed_dim != 60 or growth_rate != 30: _unsupported(f"embed_dim={embed_dim}/growth_rate={growth_rate}")
        if any(w != WINDOW for w in window_size): _unsupported(f"window_size={window_size}")
        if any(h != 6 for h in num_heads): _unsupported(f"num_heads={num_heads}")
        if any(d != 2 for d in dense_layer_depths): _unsupported(f"dense_layer_depths={dense_layer_depths}")
        if any(d != 3 for d in rdb_depths): _unsupported(f"rdb_depths={rdb_depths}")
        if norm_layer is not nn.LayerNorm: _unsupported("norm_layer other than nn.LayerNorm")
        if ape: _unsupported("absolute position embedding")
        if not patch_norm: _unsupported("patch_norm=False")
        if resi_connection not in ('1conv', '3conv'): _unsupported(f"resi_connection={resi_connection!r}")
        if dim_modify_mode not in ('tail', 'head') or not pre_norm: _unsupported("dim_modify_mode other than 'tail' / 'head', or pre_norm=False")
        if scale_free or scale_embedding: _unsupported("scale_free / scale_embedding")
        if int(sr_scale) not in (2, 4): _unsupported(f"sr_scale={sr_scale}")
        if bn_in_conv: _unsupported("bn_in_conv")
        if drop_rate or attn_drop: _unsupported("dropout > 0")
        if not qkv_bias: _unsupported("qkv_bias=False")
        # the fused MLP kernels and their packed weight images are built for hidden = 2C (Hp = 128 / 192 / 240)
        if float(mlp_ratio) != 2.0: _unsupported(f"mlp_ratio={mlp_ratio} (the E1 family uses swin_hidden_ratio = 2)")

        img = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        self.input_resolution, self.patch_size, self.input_channel = img_size, patch_size, in_chans
        self.num_blocks, self.n_feats = n, embed_dim
        self.sr_scale = int(sr_scale)
        self.mlp_ratio, self.growth_rate = mlp_ratio, growth_rate
        self.dense_scale, self.rdb_residual_scale, self.global_res_scale = dense_scale, rdb_residual_scale, global_res_scale
        self.feature_last_operation = feature_last_operation
        self.drop_path_rate = drop_path_rate      # stored, never applied (as in the reference, :1172)
        self.use_checkpoint = use_checkpoint

        mean = [0.] * in_chans if mean is None else list(mean)
        std = [1.] * in_chans if std is None else list(std)
        if len(mean) != len(std) or len(mean) != in_chans:
            raise ValueError('Dimension of mean {} / std {} should fit input channels {}'.format(len(mean), len(std), in_chans))
        self.mean, self.std = mean, std
        # registration order follows the reference so state_dict() iterates identically
        self.add_mean = MeanShift(mean, std, 'add')
        self.sub_mean = MeanShift(mean, std, 'sub')
        self.head = nn.Conv2d(in_chans, embed_dim, 3, padding=1)
        self.patch_embed = PatchEmbed(embed_dim, patch_norm)
        self.body = nn.ModuleList([
            RDSTB(embed_dim, img, dense_layer_depths[i], num_heads[i], mlp_ratio, qkv_bias, qk_scale,
                  growth_rate, rdb_depths[i], resi_connection, dim_modify_mode) for i in range(n)])
        self.resi_connection, self.dim_modify_mode = resi_connection, dim_modify_mode
        self.norm = nn.LayerNorm(embed_dim)
        if resi_connection == '1conv':
            self.conv_after_body = nn.Conv2d(embed_dim, embed_dim, 3, 1, 1)
        else:                                       # '3conv' (rdst_variations.py:1286-1292)
            self.conv_after_body = nn.Sequential(nn.Conv2d(embed_dim, embed_dim // 4, 3, 1, 1), nn.LeakyReLU(negative_slope=0.2, inplace=True),
                                                 nn.Conv2d(embed_dim // 4, embed_dim // 4, 1, 1, 0), nn.LeakyReLU(negative_slope=0.2, inplace=True),
                                                 nn.Conv2d(embed_dim // 4, embed_dim, 3, 1, 1))
        up = []
        s = self.sr_scale
        while s > 1:
            up += [nn.Conv2d(embed_dim, 4 * embed_dim, 3, padding=1), nn.PixelShuffle(2)]
            s //= 2
        self.tail = nn.Sequential(nn.Sequential(*up), nn.Conv2d(embed_dim, in_chans, 3, padding=1))
        self.apply(self._init_weights)

        self.precision = precision or os.environ.get("RDST_B200_PRECISION", "fp32")
        self._exec = executor.Executor(self)

    @staticmethod
    def _init_weights(m):                      # reference :1309-1316
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02, a=-2., b=2.)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    # -- precision of the CUDA path: 'fp32' (FFMA, <=1e-4 of the reference) or 'bf16' (tcgen05 tensor cores) --
    def set_precision(self, precision):
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision = precision
        return self

    def forward(self, x, sr_scale=None, out=None):
        """Reference signature forward(x, sr_scale=None) (rdst_variations.py:1342); `out` is an optional extension: an fp32
        result tensor on the device or in pinned host memory that the last kernel writes directly (inference only)."""
        if not self._exec.bound_to(self):            # e.g. after copy.deepcopy
            self._exec = executor.Executor(self)
        return self._exec.forward(x, out)

    def extra_repr(self):
        return f"precision={self.precision}, sr_scale={self.sr_scale}, backend=librdst_b200 (sm_100a)"

    def flops(self):
        return None


class RDSTSR_N(RDSTSR):
    """Same constructor signature as the reference RDSTSR_N (rdst_variations.py:850-867).  Supported: the E1 envelope of
    RDSTSR plus global_bottleneck=True, global_bottleneck_ratio=1, global_bottleneck_mode 'mlp' (cat of all RDSTB outputs
    -> Linear(60n, 60) -> Linear(60, 60), :995-1002, :1071-1079) or 'conv' (-> 1x1 conv 60n->60 -> 3x3 conv 60->60, :1003-1007, :1080-1082).  As in the reference, `norm` and `conv_after_body` are
    registered (they are in the state_dict) but not used by the forward, so they receive no gradient."""

    def __init__(self, img_size=48, patch_size=1, in_chans=1, sr_scale=2, embed_dim=60,
                 dense_layer_depths=[2, 2, 2, 2], num_heads=[6, 6, 6, 6],
                 window_size=[4, 4, 4, 4], rdb_depths=[3, 3, 3, 3],
                 mlp_ratio=4., qkv_bias=True, qk_scale=None,
                 drop_rate=0., attn_drop=0., drop_path_rate=0.,
                 norm_layer=nn.LayerNorm, ape=False, patch_norm=True,
                 use_checkpoint=False, resi_connection='1conv',
                 growth_rate=30, dense_scale=1., dim_modify_mode='tail',
                 rdb_residual_scale=1., global_res_scale=1.,
                 mean=None, std=None,
                 act_in_conv='leaky_relu', bn_in_conv=None,
                 scale_free=False, scale_embedding=False,
                 pre_norm=False,
                 global_bottleneck=True, global_bottleneck_ratio=1., global_bottleneck_mode='mlp',
                 precision=None):
        super().__init__(img_size=img_size, patch_size=patch_size, in_chans=in_chans, sr_scale=sr_scale, embed_dim=embed_dim,
                         dense_layer_depths=dense_layer_depths, num_heads=num_heads, window_size=window_size,
                         rdb_depths=rdb_depths, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                         drop_rate=drop_rate, attn_drop=attn_drop, drop_path_rate=drop_path_rate, norm_layer=norm_layer,
                         ape=ape, patch_norm=patch_norm, use_checkpoint=use_checkpoint, resi_connection=resi_connection,
                         growth_rate=growth_rate, dense_scale=dense_scale, dim_modify_mode=dim_modify_mode,
                         rdb_residual_scale=rdb_residual_scale, global_res_scale=global_res_scale, mean=mean, std=std,
                         act_in_conv=act_in_conv, bn_in_conv=bn_in_conv, scale_free=scale_free,
                         scale_embedding=scale_embedding, pre_norm=pre_norm, feature_last_operation=False,
                         precision=precision)
        if not global_bottleneck: _unsupported("RDSTSR_N with global_bottleneck=False")
        if float(global_bottleneck_ratio) != 1.0: _unsupported(f"global_bottleneck_ratio={global_bottleneck_ratio}")
        if global_bottleneck_mode not in ('mlp', 'conv'): _unsupported(f"global_bottleneck_mode={global_bottleneck_mode!r}")
        self.global_bottleneck_mode, self.do_global_bottleneck = global_bottleneck_mode, True
        del self.feature_last_operation                       # not an attribute of the reference RDSTSR_N
        if global_bottleneck_mode == 'mlp':                   # :998-1002
            self.bottleneck = nn.Sequential(nn.Linear(embed_dim * self.num_blocks, embed_dim), nn.Linear(embed_dim, embed_dim))
        else:                                                 # :1003-1007: 1x1 conv, then 3x3 conv (default_conv: same padding)
            self.bottleneck = nn.Sequential(nn.Conv2d(embed_dim * self.num_blocks, embed_dim, 1),
                                            nn.Conv2d(embed_dim, embed_dim, 3, padding=1))
        self.bottleneck.apply(self._init_weights)
        # registration order of the reference: ... body, norm, bottleneck, conv_after_body, tail
        order = list(self._modules)
        order.remove("bottleneck")
        order.insert(order.index("norm") + 1, "bottleneck")
        mods = dict(self._modules)
        self._modules.clear()
        for k in order:
            self._modules[k] = mods[k]
        self._exec = executor.ExecutorN(self)

    def forward(self, x, sr_scale=None, out=None):
        if not self._exec.bound_to(self):
            self._exec = executor.ExecutorN(self)
        return self._exec.forward(x, out)


class RRDSTB(nn.Module):
    """Parameter container of rdst_variations.py:464-555: RDSTBs + 3x3 conv ('1conv'), * rrdb_residual_scale, + input."""

    def __init__(self, input_dim, input_resolution, layer_depth, num_heads, mlp_ratio, qkv_bias, qk_scale, growth_rate,
                 num_blocks_in_rdb, num_blocks_in_rrdb, rrdb_residual_scale):
        super().__init__()
        self.residual_scale = rrdb_residual_scale
        self.body = nn.ModuleList([RDSTB(input_dim, input_resolution, layer_depth, num_heads, mlp_ratio, qkv_bias, qk_scale,
                                         growth_rate, num_blocks_in_rdb) for _ in range(int(num_blocks_in_rrdb))])
        self.conv = nn.Conv2d(input_dim, input_dim, 3, 1, 1)


class ESTSR(RDSTSR):
    """Same constructor signature as the reference ESTSR (rdst_variations.py:602-619): residual-in-residual RDSTBs.
    body.i is an RRDSTB (rrdb_depths[i] RDSTBs + conv); the forward is head -> RRDSTBs -> norm * global_res_scale + head
    output -> tail (:783-812) -- `conv_after_body` is registered, as in the reference, but not used.  The reference has
    no factory for this class; it is exported as rdst_b200.ESTSR."""

    def __init__(self, img_size=48, patch_size=1, in_chans=1, sr_scale=2, embed_dim=60,
                 dense_layer_depths=[2, 2, 2, 2], num_heads=[6, 6, 6, 6],
                 window_size=[4, 4, 4, 4], rdb_depths=[3, 3, 3, 3],
                 rrdb_depths=[3, 3, 3, 3], num_rrdb_blocks=4,
                 mlp_ratio=4., qkv_bias=True, qk_scale=None,
                 drop_rate=0., attn_drop=0., drop_path_rate=0.,
                 norm_layer=nn.LayerNorm, ape=False, patch_norm=True,
                 use_checkpoint=False, resi_connection='1conv',
                 growth_rate=30, dense_scale=1., dim_modify_mode='tail',
                 rdb_residual_scale=1., rrdb_residual_scale=1., global_res_scale=1.,
                 mean=None, std=None,
                 act_in_conv='leaky_relu', bn_in_conv=None,
                 scale_free=False, pre_norm=False, precision=None):
        n = int(num_rrdb_blocks)
        if not (len(dense_layer_depths) >= n and len(num_heads) >= n and len(window_size) >= n and len(rdb_depths) >= n and
                len(rrdb_depths) >= n):
            raise AssertionError("per-RRDSTB lists are shorter than num_rrdb_blocks")
        super().__init__(img_size=img_size, patch_size=patch_size, in_chans=in_chans, sr_scale=sr_scale, embed_dim=embed_dim,
                         dense_layer_depths=list(dense_layer_depths)[:n], num_heads=list(num_heads)[:n],
                         window_size=list(window_size)[:n], rdb_depths=list(rdb_depths)[:n], mlp_ratio=mlp_ratio,
                         qkv_bias=qkv_bias, qk_scale=qk_scale, drop_rate=drop_rate, attn_drop=attn_drop,
                         drop_path_rate=drop_path_rate, norm_layer=norm_layer, ape=ape, patch_norm=patch_norm,
                         use_checkpoint=use_checkpoint, resi_connection=resi_connection, growth_rate=growth_rate,
                         dense_scale=dense_scale, dim_modify_mode=dim_modify_mode, rdb_residual_scale=rdb_residual_scale,
                         global_res_scale=global_res_scale, mean=mean, std=std, act_in_conv=act_in_conv,
                         bn_in_conv=bn_in_conv, scale_free=scale_free, pre_norm=pre_norm, feature_last_operation=False,
                         precision=precision)
        if resi_connection != '1conv' or dim_modify_mode != 'tail':
            _unsupported("ESTSR with resi_connection other than '1conv' or dim_modify_mode other than 'tail'")
        img = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        self.body = nn.ModuleList([
            RRDSTB(embed_dim, img, dense_layer_depths[i], num_heads[i], mlp_ratio, qkv_bias, qk_scale, growth_rate,
                   rdb_depths[i], rrdb_depths[i], rrdb_residual_scale) for i in range(n)])
        self.body.apply(self._init_weights)
        self.rrdb_residual_scale = rrdb_residual_scale
        self._exec = executor.ExecutorE(self)

    def forward(self, x, sr_scale=None, out=None):
        if not self._exec.bound_to(self):
            self._exec = executor.ExecutorE(self)
        return self._exec.forward(x, out)


def make_RDSTSR(paras, mean=None, std=None):
    """Same contract as the reference factory (rdst_variations.py:1369-1457): reads the same `paras.*` names."""
    norm_layer = nn.LayerNorm if paras.rdst_layer_norm else nn.Identity
    if paras.rdst_global_bottleneck:
        return RDSTSR_N(
            img_size=paras.patch_size, patch_size=paras.swin_patch_size, in_chans=paras.input_channel,
            sr_scale=int(paras.sr_scale), embed_dim=paras.rdst_embed_dim,
            dense_layer_depths=paras.rdst_dense_layer_depths, num_heads=paras.rdst_num_heads,
            window_size=paras.rdst_window_size, rdb_depths=paras.rdst_rdb_depths,
            mlp_ratio=paras.swin_hidden_ratio, qkv_bias=paras.swin_qkv_bias, qk_scale=paras.swin_qk_scale,
            drop_rate=paras.swin_drop_rate, attn_drop=paras.swin_attn_drop_rate,
            drop_path_rate=paras.swin_drop_path_rate,
            norm_layer=norm_layer, ape=paras.rdst_ape, patch_norm=paras.rdst_patch_norm,
            use_checkpoint=paras.rdst_use_checkpoint, resi_connection=paras.rdst_res_connection,
            growth_rate=paras.rdst_growth_rate, dense_scale=paras.rdst_dense_scale,
            dim_modify_mode=paras.rdst_dim_modify_mode,
            rdb_residual_scale=paras.rdst_rdb_residual_scale, global_res_scale=paras.rdst_global_res_scale,
            mean=mean, std=std, act_in_conv=paras.rdst_act_in_conv, bn_in_conv=paras.rdst_bn_in_conv,
            scale_free=paras.scale_free, pre_norm=paras.rdst_pre_norm,
            global_bottleneck=True, global_bottleneck_ratio=paras.rdst_global_bottleneck_ratio,
            global_bottleneck_mode=paras.rdst_global_bottleneck_mode,
            precision=getattr(paras, "rdst_b200_precision", None))
    _ = paras.rdst_global_bottleneck_ratio            # read for interface parity; unused on this branch
    return RDSTSR(
        img_size=paras.patch_size, patch_size=paras.swin_patch_size, in_chans=paras.input_channel,
        sr_scale=int(paras.sr_scale), embed_dim=paras.rdst_embed_dim,
        dense_layer_depths=paras.rdst_dense_layer_depths, num_heads=paras.rdst_num_heads,
        window_size=paras.rdst_window_size, rdb_depths=paras.rdst_rdb_depths,
        mlp_ratio=paras.swin_hidden_ratio, qkv_bias=paras.swin_qkv_bias, qk_scale=paras.swin_qk_scale,
        drop_rate=paras.swin_drop_rate, attn_drop=paras.swin_attn_drop_rate,
        drop_path_rate=paras.swin_drop_path_rate,
        norm_layer=norm_layer, ape=paras.rdst_ape, patch_norm=paras.rdst_patch_norm,
        use_checkpoint=paras.rdst_use_checkpoint, resi_connection=paras.rdst_res_connection,
        growth_rate=paras.rdst_growth_rate, dense_scale=paras.rdst_dense_scale,
        dim_modify_mode=paras.rdst_dim_modify_mode,
        rdb_residual_scale=paras.rdst_rdb_residual_scale, global_res_scale=paras.rdst_global_res_scale,
        mean=mean, std=std,
        act_in_conv=paras.rdst_act_in_conv, bn_in_conv=paras.rdst_bn_in_conv,
        scale_free=paras.scale_free,
        pre_norm=paras.rdst_pre_norm,
        feature_last_operation=paras.rdst_feature_last_operation,
        precision=getattr(paras, "rdst_b200_precision", None),
    )
