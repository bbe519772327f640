"""Host-side weight packing for the librdst_b200 kernels.

The kernels see activations in a token-major, padded channel layout (include/rdst_b200.h) and expect
  * LayerNorm affine folded into the following Linear        (W' = W * gamma, b' = b + W @ beta),
  * the attention scale head_dim**-0.5 folded into the q rows of qkv (reference scales q before q@k^T,
    networks/swin_transformer_sr.py:120),
  * K columns / N rows placed at padded channel positions with zeros at the pads,
  * conv weights as [N][tap][Cin_padded] and the PixelShuffle permutation folded into the row order of
    the up-sampling convs (networks/common.py:129-132).
Everything here is tiny (4.5 M parameters) and runs as ordinary torch ops on the parameter's device; the
result is a cache keyed on parameter versions, never part of state_dict.
"""
import torch

EMBED = 60
GROWTH = 30
HEADS = 6
DENSE_LD = 160          # 64 (trunk, 60 real) + 3 * 32 (growth, 30 real)
FEAT_LD = 64


_DIFFERENTIABLE = False


class differentiable:
    """Context manager: packing keeps the autograd graph to the parameters (training path)."""

    def __enter__(self):
        global _DIFFERENTIABLE
        self._old, _DIFFERENTIABLE = _DIFFERENTIABLE, True

    def __exit__(self, *exc):
        global _DIFFERENTIABLE
        _DIFFERENTIABLE = self._old


def _f(t):
    return t.float() if _DIFFERENTIABLE else t.detach().float()


def padded_width(c):
    """Stored width of a dense-block activation with c = 60 + 30*j real channels (c = 30: the Swin blocks of a
    'head'-mode DenseSTLayer run at the growth width, stored 32 wide)."""
    if c == GROWTH:
        return 32
    j = (c - EMBED) // GROWTH
    assert c == EMBED + GROWTH * j and 0 <= j <= 3, f"unsupported channel count {c}"
    return 64 + 32 * j


def channel_positions(c, device=None):
    """Stored position of every real channel: trunk at [0,60), growth block g at [64+32g, +30)."""
    idx = torch.arange(c, device=device)
    if c == GROWTH:
        return idx
    g = torch.clamp((idx - EMBED) // GROWTH, min=0)
    return torch.where(idx < EMBED, idx, 64 + 32 * g + (idx - EMBED) % GROWTH)


def hidden_width(h):
    return (h + 15) // 16 * 16


def scatter_cols(w, pos, width):
    """[N][C] -> [N][width] with column c moved to pos[c]."""
    out = w.new_zeros(w.shape[0], width)
    out[:, pos] = w
    return out


def scatter_rows(w, pos, height):
    out = w.new_zeros((height,) + tuple(w.shape[1:]))
    out[pos] = w
    return out


def fold_ln(w, b, gamma, beta):
    """Linear(LN(x)) with affine == Linear'(xhat):  W' = W*gamma,  b' = b + W@beta."""
    return w * gamma[None, :], b + w @ beta


def pack_stl(blk, c):
    """blk: Swin block parameter container (norm1, attn.{qkv,proj,relative_position_bias_table}, norm2, mlp)."""
    f = _f
    cp = padded_width(c)
    pos = channel_positions(c, blk.norm1.weight.device)
    hd = c // HEADS
    wq, bq = fold_ln(f(blk.attn.qkv.weight), f(blk.attn.qkv.bias), f(blk.norm1.weight), f(blk.norm1.bias))
    scale = blk.attn.scale
    wq = wq.clone(); bq = bq.clone()
    wq[:c] *= scale
    bq[:c] *= scale
    hid = blk.mlp.fc1.weight.shape[0]
    hp = hidden_width(hid)
    w1, b1 = fold_ln(f(blk.mlp.fc1.weight), f(blk.mlp.fc1.bias), f(blk.norm2.weight), f(blk.norm2.bias))
    w1p = w1.new_zeros(hp, cp); w1p[:hid] = scatter_cols(w1, pos, cp)
    b1p = b1.new_zeros(hp); b1p[:hid] = b1
    w2 = f(blk.mlp.fc2.weight)                     # [C][hid]
    w2p = w2.new_zeros(cp, hp); w2p[pos, :hid] = w2
    return dict(
        c=c, cp=cp, hd=hd, hp=hp,
        wqkv=scatter_cols(wq, pos, cp).contiguous(), bqkv=bq.contiguous(),                    # [3C][Cp], [3C]
        wproj=scatter_rows(f(blk.attn.proj.weight), pos, cp).contiguous(),                    # [Cp][C]
        bproj=scatter_rows(f(blk.attn.proj.bias), pos, cp).contiguous(),
        w1=w1p.contiguous(), b1=b1p.contiguous(), w2=w2p.contiguous(),
        b2=scatter_rows(f(blk.mlp.fc2.bias), pos, cp).contiguous(),
        table=f(blk.attn.relative_position_bias_table).contiguous(),                           # [225][heads]
    )


def pack_dstl_tail(dstl, c, dense_scale):
    """LN(C) -> Linear(C, growth), written as a 32-wide slice (30 real) of the dense buffer."""
    f = _f
    cp = padded_width(c)
    pos = channel_positions(c, dstl.tail[0].weight.device)
    w, b = fold_ln(f(dstl.tail[1].weight), f(dstl.tail[1].bias), f(dstl.tail[0].weight), f(dstl.tail[0].bias))
    g = w.shape[0]
    wp = w.new_zeros(32, cp); wp[:g] = scatter_cols(w, pos, cp)
    bp = b.new_zeros(32); bp[:g] = b
    return dict(w=wp.contiguous(), b=bp.contiguous(), scale=float(dense_scale), wimg=kmajor_image(wp))


def pack_dstl_head(dstl, c):
    """'head' mode (rdst_variations.py:288-295): LN(C) -> Linear(C, growth) in FRONT of the Swin blocks; output stored 32 wide."""
    f = _f
    cp = padded_width(c)
    pos = channel_positions(c, dstl.head[0].weight.device)
    w, b = fold_ln(f(dstl.head[1].weight), f(dstl.head[1].bias), f(dstl.head[0].weight), f(dstl.head[0].bias))
    g = w.shape[0]
    wp = w.new_zeros(32, cp); wp[:g] = scatter_cols(w, pos, cp)
    bp = b.new_zeros(32); bp[:g] = b
    return dict(w=wp.contiguous(), b=bp.contiguous())


def pack_conv(weight, bias, cin_pos, cin_width, n_pad):
    """Conv2d weight (N, Cin, 3, 3) -> [n_pad][9][cin_width] (tap = ky*3+kx), bias -> [n_pad]."""
    w = _f(weight)
    n, cin = w.shape[0], w.shape[1]
    wt = w.permute(0, 2, 3, 1).reshape(n, 9, cin)
    out = w.new_zeros(n_pad, 9, cin_width)
    out[:n, :, cin_pos] = wt
    b = w.new_zeros(n_pad)
    b[:n] = _f(bias)
    return out.contiguous(), b.contiguous()


def pack_upconv(weight, bias, group=FEAT_LD):
    """UpSampler conv (4*F, F, 3, 3) + PixelShuffle(2): row order becomes (sub-pixel s = 2*dy+dx, channel c),
    s-major, each group padded to `group` rows; reference out-channel index is c*4 + s."""
    w = _f(weight)
    f4, fin = w.shape[0], w.shape[1]
    fo = f4 // 4
    wt = w.permute(0, 2, 3, 1).reshape(fo, 4, 9, fin)          # [c][s][tap][cin]
    out = w.new_zeros(4, group, 9, FEAT_LD)
    out[:, :fo, :, :fin] = wt.permute(1, 0, 2, 3)
    b = w.new_zeros(4, group)
    b[:, :fo] = _f(bias).reshape(fo, 4).t()
    return out.reshape(4 * group, 9, FEAT_LD).contiguous(), b.reshape(-1).contiguous()


# ---------------------------------------------------------------------------------------------------------
# tcgen05 operand images (bf16).  A K-major SWIZZLE_NONE operand with R rows and K columns is stored as
# [K/8][R][8]: 8x16-byte core matrices, consecutive rows 16 B apart, K-chunks R*16 B apart (rdst_b200/csrc/umma.cuh).
# ---------------------------------------------------------------------------------------------------------
def kmajor_image(w, dtype=torch.bfloat16):
    """[R][K] fp32 -> [K/8][R][8] 16-bit (flat)."""
    r, k = w.shape
    assert k % 8 == 0
    return w.to(dtype).reshape(r, k // 8, 8).permute(1, 0, 2).contiguous().reshape(-1)


def fc1_image(w1):
    """fc1 weight image: fp16 (fc1 runs on fp16 operands; the warp-specialised kernel accumulates it in fp16 so that the GELU
    gets its pre-activations as packed pairs straight out of TMEM)."""
    return kmajor_image(w1, torch.float16)


def fc2_image(w2):
    """fc2 weight image: fp16 (the GELU hidden activations are kept in fp16, see tc_mlp.cu), scaled by 1/2: the kernels'
    GELU stage emits x*(1 + tanh(..)) = 2*GELU(x) (one instruction less per pair; the scaling is exact)."""
    return kmajor_image(w2 * 0.5, torch.float16)


LOG2E = 1.4426950408889634


def attn_shapes(c):
    hd = c // HEADS
    nh = (3 * hd + 15) // 16 * 16
    hdo = 20 if hd == 20 else 16
    kproj = (HEADS * hdo + 15) // 16 * 16
    return hd, nh, hdo, kproj


def pack_attn_tc(wqkv, bqkv, wproj, bproj, table, c):
    """wqkv [3C][Cp] / bqkv [3C] (LN-folded, q rows scaled), wproj [Cp][C], bproj [Cp], table [225][heads]
    -> operand images of rdst_stl_attn_fwd_bf16 (include/rdst_b200.h)."""
    hd, nh, hdo, kproj = attn_shapes(c)
    cp = wqkv.shape[1]
    imgs, biases = [], []
    for h in range(HEADS):
        wh = wqkv.new_zeros(nh, cp)
        bh = bqkv.new_zeros(nh)
        for s in range(3):                                   # q | k | v rows of this head
            rows = slice(s * c + h * hd, s * c + (h + 1) * hd)
            f = LOG2E if s == 0 else 1.0                     # softmax is evaluated with exp2
            wh[s * hd:(s + 1) * hd] = wqkv[rows] * f
            bh[s * hd:(s + 1) * hd] = bqkv[rows] * f
        imgs.append(kmajor_image(wh))
        biases.append(bh)
    wp = wproj.new_zeros(cp, kproj)
    for h in range(HEADS):
        wp[:, h * hdo:h * hdo + hd] = wproj[:, h * hd:(h + 1) * hd]
    # relative-position bias, pre-scaled by log2(e), as packed fp16 pairs (t[dy][dx], t[dy][dx-1]): the kernel adds the
    # bias of two neighbouring keys with one 32-bit shared-memory read (row pitch 24 -> bank-conflict free)
    t15 = (table.t() * LOG2E).reshape(HEADS, 15, 15)
    lo = table.new_zeros(HEADS, 15, 24)
    hi = table.new_zeros(HEADS, 15, 24)
    lo[:, :, :15] = t15
    hi[:, :, 1:15] = t15[:, :, :14]
    tab = torch.stack([lo, hi], dim=-1).to(torch.float16).contiguous().view(torch.int32).reshape(HEADS, 15, 24)
    return dict(wqkv_img=torch.cat(imgs).contiguous(), bqkv_tc=torch.cat(biases).contiguous(),
                wproj_img=kmajor_image(wp), table_tc=tab.contiguous())


def pack_stl_tc(p):
    """Add tensor-core operand images to a pack_stl() dict."""
    p["w1img"] = fc1_image(p["w1"])
    p["w2img"] = fc2_image(p["w2"])
    p.update(pack_attn_tc(p["wqkv"], p["bqkv"], p["wproj"], p["bproj"], p["table"], p["c"]))
    return p


def conv_tc_image(w):
    """[N][9][Cin] fp32 conv weight -> operand images of rdst_conv3x3_fwd_bf16_tc: N/NT slices x 9 taps x [Cin/8][NT][8]."""
    n, _, cin = w.shape
    nt = 32 if cin == 160 else (128 if n == 256 else 64)
    parts = [kmajor_image(w[s * nt:(s + 1) * nt, tap, :]) for s in range(n // nt) for tap in range(9)]
    return torch.cat(parts).contiguous()


def last_conv_tc_image(w9):
    """[9][64] fp32 filter of the final 64->1 conv -> one K-major image [8][16][8] bf16 whose rows are the 9 taps
    (rows 9..15 zero): B operand of the tap GEMM P[pos][tap] = x[pos] . w[tap] (tc_conv.cu, last_conv_tap_kernel)."""
    m = w9.new_zeros(16, w9.shape[1])
    m[:9] = w9
    return kmajor_image(m).contiguous()
