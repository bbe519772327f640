"""Training path: forward with saved activations + hand-written backward over librdst_b200 kernels (fp32).

The network is a chain of torch.autograd.Functions (head, one per RDSTB, tail).  Their tensor inputs are the *packed*
weights (rdst_b200/packing.py run in differentiable mode: LayerNorm folding, q scaling, padded scatter are ordinary
torch ops on the 4.5 M parameters), so autograd carries the gradients of the packed tensors back to the reference-named
parameters; every activation-sized computation, forward and backward, is a kernel of the C ABI:
    data gradients     rdst_linear_fwd / rdst_conv3x3_fwd with transposed (and tap-flipped) weights
    weight gradients   rdst_gemm_tn_acc (token reduction, optional 3x3 gather) incl. bias column sums
    LayerNorm          rdst_lnhat_fwd/_bwd (affine folded), rdst_layernorm_bwd (patch_embed.norm, final norm)
    attention          rdst_window_attention_bwd (recomputes the probabilities; relative-position table gradient)
    GELU, residual joins, PixelShuffle  rdst_gelu_fwd/_bwd, rdst_axpy, rdst_pixel_unshuffle2
Gradient of: RDSTSR.forward (rdst_variations.py:1342-1360) and everything it calls; checked against torch.autograd
through the CPU oracle in tests/test_gpu_backward.py (fp32) and tests/test_gpu_train_tc.py (bf16 tensor-core GEMMs).
"""
import contextlib
import threading

import torch

from . import _lib, packing

F32 = _lib.F32


def _call(name, *a):
    _lib.call(name, *a)


def _p(t):
    return _lib.ptr(t)


def _ld(t):
    return t.stride(0) if t.dim() == 2 else 1


# ------------------------------------------------------------------------------------------------ kernel wrappers
# precision 'fp32': every GEMM is a CUDA-core FFMA kernel (the <=1e-4 parity path).  precision 'bf16': the GEMMs --
# Linear / conv forward, data gradients, weight gradients -- run on tcgen05 with operands rounded to bf16 while they
# are staged (csrc/tc_train.cu); storage, LayerNorm / softmax / GELU and all reductions stay fp32.  Each Function sets
# the mode of the calling thread on entry (backward runs on autograd's device thread).
_MODE = threading.local()


def _tc():
    return getattr(_MODE, "tc", False)


def _aligned(*ts):
    return all(t is None or (t.data_ptr() % 16 == 0 and (t.dim() < 2 or t.stride(0) % 4 == 0)) for t in ts)


def linear(x, w, b, y, K, N, ln_creal=0, scale=1.0, resid=None, gelu_in=False):
    """y = scale * (op(x) . w^T + b) + resid;  op = LayerNorm-hat (ln_creal > 0), exact-erf GELU (gelu_in) or identity."""
    T = x.shape[0]
    if _tc() and _aligned(x, y, resid):
        _call("rdst_gemm_tc", _p(x), _ld(x), _p(w), _ld(w), 0, _p(b), _p(resid), 0 if resid is None else _ld(resid), None, 0,
              _p(y), _ld(y), T, K, N, 1 if ln_creal else (2 if gelu_in else 0), ln_creal, scale, 0, 0, 0, 0, 0, 0,
              _lib.stream_ptr())
        return
    assert w.is_contiguous() and w.shape[-1] == K
    if gelu_in:
        act = torch.empty(T, K, dtype=torch.float32, device=x.device)
        _call("rdst_gelu_fwd", _p(x), _ld(x), _p(act), K, T, K, _lib.stream_ptr())
        x = act
    _call("rdst_linear_fwd", _p(x), _ld(x), _p(w), _p(b), _p(resid), 0 if resid is None else _ld(resid), _p(y), _ld(y),
          T, K, N, ln_creal, 0, scale, F32, _lib.stream_ptr())


def linear_t(dy, w, dx, K, N, scale=1.0, gelu_aux=None):
    """Data gradient of a Linear: dx[T][N] = scale * dy[T][K] . w[K][N]   (w is the forward weight [out=K][in=N]),
    optionally times gelu'(gelu_aux) -- the gradient through the GELU that fed the Linear."""
    T = dy.shape[0]
    if _tc() and _aligned(dy, dx, gelu_aux):
        _call("rdst_gemm_tc", _p(dy), _ld(dy), _p(w), _ld(w), 1, None, None, 0, _p(gelu_aux),
              0 if gelu_aux is None else _ld(gelu_aux), _p(dx), _ld(dx), T, K, N, 0, 0, scale, 0, 0, 0, 0, 0, 0,
              _lib.stream_ptr())
        return
    wt = w.t().contiguous()
    zb = torch.zeros(N, dtype=torch.float32, device=dy.device)
    if gelu_aux is None:
        linear(dy, wt, zb, dx, K, N, scale=scale)
    else:
        tmp = torch.empty(T, N, dtype=torch.float32, device=dy.device)
        linear(dy, wt, zb, tmp, K, N, scale=scale)
        _call("rdst_gelu_bwd", _p(gelu_aux), _ld(gelu_aux), _p(tmp), N, _p(dx), _ld(dx), T, N, _lib.stream_ptr())


def linear_t_lnbwd(dy, w, x, dx, K, N, creal, scale=1.0, resid=None, resid2=None):
    """dx = lnhat_bwd(scale * dy . w, x) + resid + resid2: data gradient through LayerNorm-hat -> Linear (w = forward weight
    [out=K][in=N]).  One tensor-core kernel in bf16 mode (the LayerNorm backward is the GEMM's epilogue)."""
    if _tc() and N <= 128 and N % 16 == 0 and _aligned(dy, x, dx, resid, resid2):
        _call("rdst_gemm_tc_lnbwd", _p(dy), _ld(dy), _p(w), _ld(w), _p(x), _ld(x), _p(resid), 0 if resid is None else _ld(resid),
              _p(resid2), 0 if resid2 is None else _ld(resid2), _p(dx), _ld(dx), dy.shape[0], K, N, creal, scale,
              _lib.stream_ptr())
        return
    dxh = torch.empty(dy.shape[0], N, dtype=torch.float32, device=dy.device)
    linear_t(dy, w, dxh, K, N, scale=scale)
    lnhat_bwd(dxh, x, dx, N, creal, resid=resid, resid2=resid2)


def conv(x, w, b, y, B, H, W, cin, n, scale=1.0, shuffle=0, resid=None):
    if _tc() and cin % 16 == 0 and _aligned(x, y, resid):
        _call("rdst_gemm_tc", _p(x), _ld(x), _p(w), 9 * cin, 0, _p(b), _p(resid), 0 if resid is None else _ld(resid), None, 0,
              _p(y), _ld(y), B * H * W, 9 * cin, n, 0, 0, scale, 1, B, H, W, cin, shuffle, _lib.stream_ptr())
        return
    _call("rdst_conv3x3_fwd", _p(x), _ld(x), _p(w), _p(b), _p(resid), 0 if resid is None else _ld(resid), _p(y), _ld(y),
          B, H, W, cin, n, scale, shuffle, F32, _lib.stream_ptr())


def gemm_tn(dy, x, dw, db, N, K, conv_geom=None, x_op=0, creal=0):
    """dw[N][K] += dy^T . op(x), db += column sums of dy;  op: 0 identity, 1 LayerNorm-hat (creal), 2 exact-erf GELU."""
    T = dy.shape[0]
    geom = (0, 0, 0, 0, 0) if conv_geom is None else (1,) + tuple(conv_geom)
    if _tc() and _aligned(dy, x) and dw.data_ptr() % 16 == 0 and (conv_geom is None or conv_geom[3] % 16 == 0):
        _call("rdst_gemm_tn_tc", _p(dy), _ld(dy), _p(x), _ld(x), _p(dw), _p(db), T, N, K, *geom, x_op, creal,
              _lib.stream_ptr())
        return
    if x_op:
        tmp = torch.empty(T, K, dtype=torch.float32, device=x.device)
        if x_op == 1:
            lnhat(x, tmp, K, creal)
        else:
            _call("rdst_gelu_fwd", _p(x), _ld(x), _p(tmp), K, T, K, _lib.stream_ptr())
        x = tmp
    _call("rdst_gemm_tn_acc", _p(dy), _ld(dy), _p(x), _ld(x), _p(dw), _p(db), T, N, K, *geom, _lib.stream_ptr())


def _pad8(n):
    return (n + 7) // 8 * 8


def lnhat(x, y, K, creal):
    _call("rdst_lnhat_fwd", _p(x), _ld(x), _p(y), _ld(y), x.shape[0], K, creal, 1, _lib.stream_ptr())


def lnhat_bwd(dxh, x, dx, K, creal, resid=None, resid2=None):
    _call("rdst_lnhat_bwd", _p(dxh), _ld(dxh), _p(x), _ld(x), _p(resid), 0 if resid is None else _ld(resid),
          _p(resid2), 0 if resid2 is None else _ld(resid2), _p(dx), _ld(dx), x.shape[0], K, creal, 1, _lib.stream_ptr())


def axpy(x, y, N, alpha=1.0):
    _call("rdst_axpy", _p(x), _ld(x), _p(y), _ld(y), x.shape[0], N, alpha, _lib.stream_ptr())


def conv_dgrad_weight(w):
    """[N][9][Cin] forward filter -> [Cin][9][N] filter of the data gradient (taps flipped)."""
    return w.flip(1).permute(2, 1, 0).contiguous()


# ------------------------------------------------------------------------------------------------ packed weights
def _adder(flat):
    def add(t):
        flat.append(t.contiguous())
        return len(flat) - 1
    return add


def block_params(blk):
    """Reference-named parameters of one RDSTB's Swin blocks and DenseSTLayer tails, in the fixed order the BlockFunction
    takes them (the LFF conv goes through packing.pack_conv under autograd: a handful of torch ops)."""
    ps = []
    for dstl in blk.body:
        for b in dstl.body.blocks:
            ps += stl_params(b)
        ps += [dstl.tail[0].weight, dstl.tail[0].bias, dstl.tail[1].weight, dstl.tail[1].bias]
    return ps


_STL_PARAMS = 13     # norm1.{w,b}, qkv.{w,b}, table, proj.{w,b}, norm2.{w,b}, fc1.{w,b}, fc2.{w,b}


def stl_params(b):
    return [b.norm1.weight, b.norm1.bias, b.attn.qkv.weight, b.attn.qkv.bias, b.attn.relative_position_bias_table,
            b.attn.proj.weight, b.attn.proj.bias, b.norm2.weight, b.norm2.bias,
            b.mlp.fc1.weight, b.mlp.fc1.bias, b.mlp.fc2.weight, b.mlp.fc2.bias]


class _Layout:
    """Where every packed tensor of one link lives in its flat packed buffer and which rdst_pack_linear descriptor produces
    it.  Slot numbers index the list of packed tensors (`W` in the Functions); parameter numbers index the link's
    parameter list."""

    def __init__(self):
        self.lins, self.slots, self.tables, self.off, self.pi = [], [], [], 0, 0

    def slot(self, shape):
        self.slots.append((self.off, tuple(shape)))
        self.off += (_numel(shape) + 3) // 4 * 4              # keep every packed tensor 16-byte aligned
        return len(self.slots) - 1

    def lin(self, w, b, g, be, N, K, rows_p, ldp, srows, scols, q_rows=0, q_scale=1.0):
        sw, sb = self.slot((rows_p, ldp)), self.slot((rows_p,))
        self.lins.append(dict(w=w, b=b, g=g, be=be, N=N, K=K, ldp=ldp, srows=srows, scols=scols, q_rows=q_rows,
                              q_scale=float(q_scale), sw=sw, sb=sb))
        return sw, sb

    def take(self, n):
        r = range(self.pi, self.pi + n)
        self.pi += n
        return r

    def stl(self, b, c):
        """One Swin block of width c (its 13 parameters are next in the link's parameter list)."""
        cp = packing.padded_width(c)
        hid = b.mlp.fc1.weight.shape[0]
        hp = packing.hidden_width(hid)
        n1g, n1b, qw, qb, tab, pw, pb, n2g, n2b, f1w, f1b, f2w, f2b = self.take(_STL_PARAMS)
        st = {"c": c, "cp": cp, "hp": hp, "shift": b.shift_size}
        st["wqkv"], st["bqkv"] = self.lin(qw, qb, n1g, n1b, 3 * c, c, 3 * c, cp, 0, 1, c, b.attn.scale)
        st["wproj"], st["bproj"] = self.lin(pw, pb, None, None, c, c, cp, c, 1, 0)
        st["w1"], st["b1"] = self.lin(f1w, f1b, n2g, n2b, hid, c, hp, cp, 0, 1)
        st["w2"], st["b2"] = self.lin(f2w, f2b, None, None, c, hid, cp, hp, 1, 0)
        st["table"] = -1 - len(self.tables)               # raw parameter, used as is: W[-1-i] is appended behind the slots
        self.tables.append(tab)
        return st

    def finish(self, bs):
        n = len(self.slots)
        bs.update(lins=self.lins, slots=self.slots, tables=self.tables, packed_floats=self.off, n_params=self.pi,
                  lff_w=n, lff_b=n + 1)                   # the link's conv filter / bias follow the slots in `W`
        return bs


def block_layout(m, blk):
    """Static description of one RDSTB (3 DenseSTLayers: 2 Swin blocks + LN/Linear tail each)."""
    L = _Layout()
    bs = {"dstl": []}
    c = packing.EMBED
    for dstl in blk.body:
        stls = [L.stl(b, c) for b in dstl.body.blocks]
        tg, tb_, tw, tbias = L.take(4)
        sw, sb = L.lin(tw, tbias, tg, tb_, packing.GROWTH, c, 32, packing.padded_width(c), 0, 1)
        bs["dstl"].append({"c": c, "stl": stls, "tw": sw, "tb": sb, "scale": float(m.dense_scale)})
        c += packing.GROWTH
    bs["res_scale"] = float(m.rdb_residual_scale)
    return L.finish(bs)


def rstb_layout(layer):
    """Static description of one RSTB of the vanilla SwinIR: depth Swin blocks at C = 60 (+ the 3x3 conv, packed by torch)."""
    L = _Layout()
    return L.finish({"stl": [L.stl(b, packing.EMBED) for b in layer.residual_group.blocks]})


def rstb_params(layer):
    return [p for b in layer.residual_group.blocks for p in stl_params(b)]


def _views(buf, slots):
    return [buf[o:o + _numel(s)].view(s) for o, s in slots]


def _numel(shape):
    n = 1
    for d in shape:
        n *= d
    return n


def pack_lff(blk, device):
    with packing.differentiable():
        pos = packing.channel_positions(packing.EMBED + 3 * packing.GROWTH, device)
        return packing.pack_conv(blk.conv.weight, blk.conv.bias, pos, packing.DENSE_LD, 64)


def pack_conv64(conv, device, n_pad=64):
    """Differentiable packing of a 60 -> n conv on a [T][64] map: [n_pad][9][64], [n_pad]."""
    with packing.differentiable():
        return packing.pack_conv(conv.weight, conv.bias, torch.arange(60, device=device), 64, n_pad)


def _pack_batch(lins, P, W, G, GP, backward):
    """All Linears of one RDSTB in one rdst_pack_linear_batch launch (forward: P -> W; backward: G -> GP)."""
    descs = []
    for l in lins:
        ln = l["g"] is not None
        d = dict(W=P[l["w"]], b=P[l["b"]], gamma=P[l["g"]] if ln else None, beta=P[l["be"]] if ln else None,
                 N=l["N"], K=l["K"], ldp=l["ldp"], scatter_rows=l["srows"], scatter_cols=l["scols"], q_rows=l["q_rows"],
                 q_scale=l["q_scale"])
        if backward:
            d.update(dWp=G[l["sw"]], dbp=G[l["sb"]], dW=GP[l["w"]], db=GP[l["b"]],
                     dgamma=GP[l["g"]] if ln else None, dbeta=GP[l["be"]] if ln else None)
        else:
            d.update(Wp=W[l["sw"]], bp=W[l["sb"]])
        descs.append(d)
    _call("rdst_pack_linear_batch", _lib.pack_desc_array(descs), len(descs), 1 if backward else 0, _lib.stream_ptr())


def pack_head(m, device):
    flat = []
    add = _adder(flat)
    with packing.differentiable():
        f = packing._f
        hw = torch.zeros(64, 9, 16, device=device)
        hw[:60, :, 0] = f(m.head.weight).reshape(60, 9)
        hb = torch.zeros(64, device=device)
        hb[:60] = f(m.head.bias)
        spec = {"head_w": add(hw), "head_b": add(hb),
                "pe_g": add(f(m.patch_embed.norm.weight)), "pe_b": add(f(m.patch_embed.norm.bias))}
    return flat, spec


def pack_tail(m, device):
    flat = []
    add = _adder(flat)
    spec = {}
    with packing.differentiable():
        f = packing._f
        id60 = torch.arange(60, device=device)
        spec["norm_g"], spec["norm_b"] = add(f(m.norm.weight)), add(f(m.norm.bias))
        w, b = packing.pack_conv(m.conv_after_body.weight, m.conv_after_body.bias, id60, 64, 64)
        spec["cab_w"], spec["cab_b"] = add(w), add(b)
        spec["up"] = []
        for l in m.tail[0]:
            if isinstance(l, torch.nn.Conv2d):
                w, b = packing.pack_upconv(l.weight, l.bias)
                spec["up"].append((add(w), add(b)))
        last = m.tail[1]
        lw = torch.zeros(1, 9, 64, device=device)
        lw[0, :, :60] = f(last.weight)[0].permute(1, 2, 0).reshape(9, 60)
        spec["last_w"], spec["last_b"] = add(lw), add(f(last.bias))
    return flat, spec


def frozen_scalars(m):
    """sub_mean / add_mean are frozen 1x1 convs (requires_grad False, reference common.py:151-167): their four numbers
    are read once per buffer version (one host sync), never inside a step -> the step stays CUDA-graph capturable."""
    key = tuple((t.data_ptr(), t._version) for t in (m.sub_mean.weight, m.sub_mean.bias, m.add_mean.weight, m.add_mean.bias))
    cache = getattr(m, "_rdst_scalar_cache", None)
    if cache is None or cache[0] != key:
        vals = dict(in_scale=float(m.sub_mean.weight.detach().reshape(-1)[0]),
                    in_bias=float(m.sub_mean.bias.detach().reshape(-1)[0]),
                    out_scale=float(m.add_mean.weight.detach().reshape(-1)[0]),
                    out_bias=float(m.add_mean.bias.detach().reshape(-1)[0]))
        cache = (key, vals)
        object.__setattr__(m, "_rdst_scalar_cache", cache)
    return dict(cache[1], grs=float(m.global_res_scale), flo=bool(getattr(m, "feature_last_operation", False)),
                sr=int(m.sr_scale))


# ------------------------------------------------------------------------------------------------ the Functions
# The network is a chain of autograd.Functions -- head, one per RDSTB, tail -- and the (differentiable) weight
# packing of each link is done right before the link runs.  In backward the autograd engine therefore finishes the
# parameter gradients of RDSTB i (link backward, then its packing graph, then AccumulateGrad) before it starts
# RDSTB i-1: gradient buckets become ready block by block and a data-parallel all-reduce (torch DDP, or
# rdst_b200.ddp.BucketedAllReduce) overlaps with the rest of the backward pass.
def _on(dev):
    """Make the tensors' device current for the launches (kernels take the current device); CPU tensors only get here in the
    `-m "not gpu"` tests, where the C ABI is replaced by its contract restatements."""
    return torch.cuda.device(dev) if dev.type == "cuda" else contextlib.nullcontext()


def _f32(dev):
    return (lambda *s: torch.empty(*s, dtype=torch.float32, device=dev),
            lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev))


class HeadFunction(torch.autograd.Function):
    """sub_mean -> head conv 1->60 -> patch_embed LayerNorm (rdst_variations.py:1344-1345, swin_transformer_sr.py:515-519).
    Returns (F0 [T][64] head-conv output, X0 [T][64] normalised trunk)."""

    @staticmethod
    def forward(ctx, spec, sc, x, *W):
        dev = x.device
        B, _, H, Wd = x.shape
        T = B * H * Wd
        e, z = _f32(dev)
        _MODE.tc = sc["tc"]
        with _on(dev):
            img = z(T, 16)
            img[:, 0] = x.reshape(-1) * sc["in_scale"] + sc["in_bias"]
            F0 = e(T, 64)
            conv(img, W[spec["head_w"]], W[spec["head_b"]], F0, B, H, Wd, 16, 64)
            X0 = z(T, 64)
            _call("rdst_layernorm_fwd", _p(F0), 64, _p(W[spec["pe_g"]]), _p(W[spec["pe_b"]]), _p(X0), 64, T, 60, 1.0, F32,
                  _lib.stream_ptr())
        ctx.spec, ctx.W, ctx.geom, ctx.saved, ctx.tc = spec, W, (B, H, Wd), (img, F0), sc["tc"]
        return F0, X0

    @staticmethod
    def backward(ctx, dF0, dX0):
        spec, W = ctx.spec, ctx.W
        B, H, Wd = ctx.geom
        T = B * H * Wd
        img, F0 = ctx.saved
        dev = dX0.device
        e, z = _f32(dev)
        G = [None] * len(W)
        _MODE.tc = ctx.tc
        with _on(dev):
            dF = dF0.contiguous().clone()
            dE = z(T, 64)
            gg, gb = z(60), z(60)
            _call("rdst_layernorm_bwd", _p(dX0.contiguous()), 64, _p(F0), 64, _p(W[spec["pe_g"]]), _p(dE), 64,
                  _p(gg), _p(gb), T, 60, 1.0, _lib.stream_ptr())
            G[spec["pe_g"]], G[spec["pe_b"]] = gg, gb
            axpy(dE, dF, 60)
            ghw, ghb = z(64, 9 * 16), z(64)
            gemm_tn(dF, img, ghw, ghb, 64, 9 * 16, (B, H, Wd, 16))
            G[spec["head_w"]], G[spec["head_b"]] = ghw.reshape(64, 9, 16), ghb
        ctx.saved = None
        return (None, None, None) + tuple(G)


class BlockFunction(torch.autograd.Function):
    """One RDSTB (rdst_variations.py:380-445): 3 DenseSTLayers on the [T][160] dense buffer, LFF conv + residual."""

    @staticmethod
    def forward(ctx, bs, geom, X, lff_w, lff_b, *P):
        B, H, Wd = geom
        T = B * H * Wd
        dev = X.device
        e, z = _f32(dev)
        _MODE.tc = tc = bs["tc"]
        with _on(dev):
            # packed weights of the block: one zeroed buffer, one rdst_pack_linear_fwd per Linear (csrc/pack.cu)
            W = _views(z(bs["packed_floats"]), bs["slots"])
            _pack_batch(bs["lins"], P, W, None, None, backward=False)
            # slots | LFF filter, bias (packed by torch ops under autograd) | relative-position tables, reversed, used as
            # stored: W[-1-i] is table i
            W = W + [lff_w, lff_b] + [P[i] for i in reversed(bs["tables"])]
            D = z(T, 160)
            D[:, :64] = X
            sb = []
            for j, ds in enumerate(bs["dstl"]):
                c = ds["c"]
                src = D
                sl = []
                for st in ds["stl"]:
                    sv = _stl_forward(st, W, src, B, H, Wd, tc)
                    sl.append(sv)
                    y = sv["y"]
                    src = y
                off = 64 + 32 * j
                linear(src, W[ds["tw"]], W[ds["tb"]], D[:, off:], ds["stl"][0]["cp"], 32, ln_creal=c, scale=ds["scale"])
                sb.append(sl)
            Xn = e(T, 64)
            conv(D, W[bs["lff_w"]], W[bs["lff_b"]], Xn, B, H, Wd, 160, 64, scale=bs["res_scale"], resid=D)
        ctx.bs, ctx.geom, ctx.W, ctx.P, ctx.saved = bs, geom, W, P, (D, sb)
        return Xn

    @staticmethod
    def backward(ctx, dXn):
        bs, W, P = ctx.bs, ctx.W, ctx.P
        B, H, Wd = ctx.geom
        T = B * H * Wd
        D, sb = ctx.saved
        dev = dXn.device
        e, z = _f32(dev)
        # gradients of the packed tensors accumulate in one zeroed buffer laid out like the packed weights
        G = _views(z(bs["packed_floats"]), bs["slots"]) + [torch.zeros_like(W[bs["lff_w"]]), torch.zeros_like(W[bs["lff_b"]])] + \
            [torch.zeros_like(P[i]) for i in reversed(bs["tables"])]

        def gz(i):
            return G[i]

        _MODE.tc = bs["tc"]
        with _on(dev):
            dX = dXn.contiguous()
            dys = dX if bs["res_scale"] == 1.0 else dX * bs["res_scale"]
            gemm_tn(dys, D, gz(bs["lff_w"]), gz(bs["lff_b"]), 64, 9 * 160, (B, H, Wd, 160))
            dD = e(T, 160)
            conv(dX, conv_dgrad_weight(W[bs["lff_w"]]), z(160), dD, B, H, Wd, 64, 160, scale=bs["res_scale"])
            axpy(dX, dD, 64)                          # residual: block output = LFF(D) + D[:, :64]
            for j in range(len(bs["dstl"]) - 1, -1, -1):
                ds, sl = bs["dstl"][j], sb[j]
                c = ds["c"]
                cp = ds["stl"][0]["cp"]
                off = 64 + 32 * j
                dg = dD[:, off:off + 32]
                y1 = sl[-1]["y"]
                gw, gb = gz(ds["tw"]), gz(ds["tb"])
                gemm_tn(dg, y1, gw, gb, 32, cp, x_op=1, creal=c)
                if ds["scale"] != 1.0:
                    gw.mul_(ds["scale"]); gb.mul_(ds["scale"])
                dy_cur = e(T, cp)
                linear_t_lnbwd(dg, W[ds["tw"]], y1, dy_cur, 32, cp, c, scale=ds["scale"])
                for k in range(len(ds["stl"]) - 1, -1, -1):
                    first = k == 0
                    dy_cur = _stl_backward(ds["stl"][k], sl[k], W, gz, dy_cur, B, H, Wd,
                                           accumulate_into=dD if first else None)
            dXin = dD[:, :64].contiguous()
            # packed-weight gradients -> gradients of the reference-named parameters
            GP = [None] * len(P)
            ln_idx = [i for l in bs["lins"] if l["g"] is not None for i in (l["g"], l["be"])]
            lnbuf = z(sum(P[i].numel() for i in ln_idx))        # dgamma / dbeta are accumulated: one zero fill for all
            off = 0
            for i in ln_idx:
                GP[i] = lnbuf[off:off + P[i].numel()].view_as(P[i])
                off += P[i].numel()
            for l in bs["lins"]:
                GP[l["w"]], GP[l["b"]] = torch.empty_like(P[l["w"]]), torch.empty_like(P[l["b"]])
            _pack_batch(bs["lins"], P, None, G, GP, backward=True)
            for i, pi in enumerate(bs["tables"]):
                GP[pi] = G[-1 - i]
        ctx.saved = None
        return (None, None, dXin, G[bs["lff_w"]], G[bs["lff_b"]]) + tuple(GP)


class TailFunction(torch.autograd.Function):
    """final norm * global_res_scale -> conv_after_body + head skip -> UpSampler -> last conv -> add_mean
    (rdst_variations.py:1337-1358)."""

    @staticmethod
    def forward(ctx, spec, sc, geom, X, F0, *W):
        B, H, Wd = geom
        T = B * H * Wd
        dev = X.device
        e, z = _f32(dev)
        _MODE.tc = sc["tc"]
        with _on(dev):
            if sc.get("tail_only"):                          # RDSTSR_N: X already is the map that feeds the up-sampler
                FN, F1 = None, X.contiguous()
            else:
                FN = z(T, 64)
                _call("rdst_layernorm_fwd", _p(X), 64, _p(W[spec["norm_g"]]), _p(W[spec["norm_b"]]), _p(FN), 64, T, 60,
                      sc["grs"], F32, _lib.stream_ptr())
                F1 = e(T, 64)
                if sc["flo"]:
                    conv(FN, W[spec["cab_w"]], W[spec["cab_b"]], F1, B, H, Wd, 64, 64, resid=F0)
                else:
                    torch.add(FN, F0, out=F1)
            feats = [F1]
            h, w_ = H, Wd
            for wi, bi in spec["up"]:
                up = e(B * 4 * h * w_, 64)
                conv(feats[-1], W[wi], W[bi], up, B, h, w_, 64, 256, shuffle=2)
                feats.append(up)
                h, w_ = 2 * h, 2 * w_
            out4 = e(B * h * w_, 4)                         # one output channel in a 16-byte row (aligned for the GEMM kernels)
            conv(feats[-1], W[spec["last_w"]], W[spec["last_b"]], out4, B, h, w_, 64, 1)
            out = (out4[:, 0] * sc["out_scale"] + sc["out_bias"]).reshape(B, 1, h, w_)
        ctx.spec, ctx.sc, ctx.geom, ctx.W, ctx.saved = spec, sc, geom, W, (X, FN, feats)
        return out

    @staticmethod
    def backward(ctx, dout):
        spec, sc, W = ctx.spec, ctx.sc, ctx.W
        B, H, Wd = ctx.geom
        T = B * H * Wd
        X, FN, feats = ctx.saved
        dev = dout.device
        e, z = _f32(dev)
        G = [None] * len(W)

        def gz(i):
            if G[i] is None:
                G[i] = torch.zeros_like(W[i])
            return G[i]

        _MODE.tc = sc["tc"]
        with _on(dev):
            n_up = len(spec["up"])
            h, w_ = H * (2 ** n_up), Wd * (2 ** n_up)
            # ---- last conv (64 -> 1) + add_mean ----
            dy16 = z(dout.numel(), 16)                      # the single output channel as column 0 of a 16-wide map
            dy16[:, 0] = dout.to(torch.float32).reshape(-1) * sc["out_scale"]
            gw16, gb16 = z(16, 9 * 64), z(16)
            gemm_tn(dy16, feats[-1], gw16, gb16, 16, 9 * 64, (B, h, w_, 64))
            G[spec["last_w"]], G[spec["last_b"]] = gw16[:1].reshape(1, 9, 64), gb16[:1]
            wT = z(64, 9, 16); wT[:, :, 0] = conv_dgrad_weight(W[spec["last_w"]])[:, :, 0]
            dfeat = e(dy16.shape[0], 64)
            conv(dy16, wT, z(64), dfeat, B, h, w_, 16, 64)
            # ---- up-sampling convs + PixelShuffle ----
            for k in range(n_up - 1, -1, -1):
                wi, bi = spec["up"][k]
                h, w_ = h // 2, w_ // 2
                tz = B * h * w_
                dz = e(tz, 256)
                _call("rdst_pixel_unshuffle2", _p(dfeat), 64, _p(dz), 256, B, h, w_, 64, _lib.stream_ptr())
                gemm_tn(dz, feats[k], gz(wi), gz(bi), 256, 9 * 64, (B, h, w_, 64))
                dfeat = e(tz, 64)
                conv(dz, conv_dgrad_weight(W[wi]), z(64), dfeat, B, h, w_, 256, 64)
            dF1 = dfeat                                   # grad wrt F1 = cab(FN) + F0
            if sc.get("tail_only"):
                ctx.saved = None
                return (None, None, None, dF1, None) + tuple(G)
            dF0 = dF1.clone()
            dFN = e(T, 64)
            if sc["flo"]:
                gemm_tn(dF1, FN, gz(spec["cab_w"]), gz(spec["cab_b"]), 64, 9 * 64, (B, H, Wd, 64))
                conv(dF1, conv_dgrad_weight(W[spec["cab_w"]]), z(64), dFN, B, H, Wd, 64, 64)
            else:
                dFN.copy_(dF1)
            dX = z(T, 64)                                 # grad wrt the trunk (last block output), pads zero
            _call("rdst_layernorm_bwd", _p(dFN), 64, _p(X), 64, _p(W[spec["norm_g"]]), _p(dX), 64,
                  _p(gz(spec["norm_g"])), _p(gz(spec["norm_b"])), T, 60, sc["grs"], _lib.stream_ptr())
        ctx.saved = None
        return (None, None, None, dX, dF0) + tuple(G)


def _stl_forward(st, W, src, B, H, Wd, tc):
    """One Swin block forward on the training path: x1 = src + proj(attn(lnhat(src))), y = x1 + fc2(gelu(fc1(lnhat(x1)))).
    Returns the tensors the backward needs (reference SwinTransformerBlock.forward, swin_transformer_sr.py:234-274)."""
    T = src.shape[0]
    c, cp, hp = st["c"], st["cp"], st["hp"]
    e, _ = _f32(src.device)
    # tensor-core mode keeps every row 16-byte aligned: q|k|v rows padded to a multiple of 8 floats
    qkv, o = e(T, _pad8(3 * c) if tc else 3 * c), e(T, cp if tc else c)
    x1, hid, y = e(T, cp), e(T, hp), e(T, cp)
    linear(src, W[st["wqkv"]], W[st["bqkv"]], qkv, cp, 3 * c, ln_creal=c)
    lse = None
    if tc:      # tcgen05 attention; keeps the row log-sum-exp for the backward kernel
        lse = e(T, packing.HEADS)
        _call("rdst_window_attention_tc_fwd", _p(qkv), _ld(qkv), _p(W[st["table"]]), _p(o), _ld(o), _p(lse),
              B, H, Wd, c, st["shift"], _lib.stream_ptr())
    else:
        _call("rdst_window_attention_fwd", _p(qkv), _ld(qkv), _p(W[st["table"]]), _p(o), _ld(o), B, H, Wd, c,
              packing.HEADS, st["shift"], F32, _lib.stream_ptr())
    linear(o, W[st["wproj"]], W[st["bproj"]], x1, c, cp, resid=src)
    linear(x1, W[st["w1"]], W[st["b1"]], hid, cp, hp, ln_creal=c)
    linear(hid, W[st["w2"]], W[st["b2"]], y, hp, cp, resid=x1, gelu_in=True)   # act = GELU(hid) not stored
    return dict(x=src, qkv=qkv, o=o, x1=x1, hid=hid, y=y, lse=lse)


def _stl_backward(st, sv, W, gz, dY, B, H, Wd, accumulate_into=None):
    """Gradient of one Swin block.  dY: [T][cp] grad of the block output.  Returns the grad of the block input, or
    (accumulate_into given) adds it in place to the first cp columns of the dense-buffer gradient and returns None."""
    T = dY.shape[0]
    c, cp, hp = st["c"], st["cp"], st["hp"]
    dev = dY.device
    e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
    x, qkv, o, x1, hid = sv["x"], sv["qkv"], sv["o"], sv["x1"], sv["hid"]
    # ---- y = x1 + fc2(gelu(fc1(lnhat(x1)))) ----
    gemm_tn(dY, hid, gz(st["w2"]), gz(st["b2"]), cp, hp, x_op=2)                     # operand GELU(hid) recomputed
    dhid = e(T, hp)
    linear_t(dY, W[st["w2"]], dhid, cp, hp, gelu_aux=hid)                            # (dY . W2) * gelu'(hid)
    gemm_tn(dhid, x1, gz(st["w1"]), gz(st["b1"]), hp, cp, x_op=1, creal=c)           # operand lnhat(x1) recomputed
    dX1 = e(T, cp)
    linear_t_lnbwd(dhid, W[st["w1"]], x1, dX1, hp, cp, c, resid=dY)
    del dhid
    # ---- x1 = x + proj(attn(lnhat(x))) ----
    gemm_tn(dX1, o, gz(st["wproj"]), gz(st["bproj"]), cp, c)
    dO = torch.empty_like(o)
    linear_t(dX1, W[st["wproj"]], dO, cp, c)
    dqkv = torch.empty_like(qkv)
    if sv["lse"] is not None:
        _call("rdst_window_attention_tc_bwd", _p(qkv), _ld(qkv), _p(W[st["table"]]), _p(sv["lse"]), _p(dO), _ld(dO),
              _p(dqkv), _ld(dqkv), _p(gz(st["table"])), B, H, Wd, c, st["shift"], _lib.stream_ptr())
    else:
        _call("rdst_window_attention_bwd", _p(qkv), _ld(qkv), _p(W[st["table"]]), _p(dO), _ld(dO), _p(dqkv), _ld(dqkv),
              _p(gz(st["table"])), B, H, Wd, c, packing.HEADS, st["shift"], _lib.stream_ptr())
    gemm_tn(dqkv, x, gz(st["wqkv"]), gz(st["bqkv"]), 3 * c, cp, x_op=1, creal=c)
    if accumulate_into is None:
        dX = e(T, cp)
        linear_t_lnbwd(dqkv, W[st["wqkv"]], x, dX, 3 * c, cp, c, resid=dX1)
        return dX
    linear_t_lnbwd(dqkv, W[st["wqkv"]], x, accumulate_into, 3 * c, cp, c, resid=dX1, resid2=accumulate_into)
    return None


# ------------------------------------------------------------------------------------------------ widened variants
def _link_packed(bs, P, conv_w, conv_b):
    """Packed tensors of a link: slots | conv filter, bias (packed by torch ops under autograd) | tables reversed."""
    dev = P[0].device
    W = _views(torch.zeros(bs["packed_floats"], dtype=torch.float32, device=dev), bs["slots"])
    _pack_batch(bs["lins"], P, W, None, None, backward=False)
    return W + [conv_w, conv_b] + [P[i] for i in reversed(bs["tables"])]


def _link_grad_buffers(bs, W, P):
    dev = P[0].device
    return _views(torch.zeros(bs["packed_floats"], dtype=torch.float32, device=dev), bs["slots"]) + \
        [torch.zeros_like(W[bs["lff_w"]]), torch.zeros_like(W[bs["lff_b"]])] + [torch.zeros_like(P[i]) for i in reversed(bs["tables"])]


def _link_param_grads(bs, P, G):
    """Packed-weight gradients -> gradients of the reference-named parameters (one rdst_pack_linear_batch launch)."""
    dev = P[0].device
    GP = [None] * len(P)
    ln_idx = [i for l in bs["lins"] if l["g"] is not None for i in (l["g"], l["be"])]
    lnbuf = torch.zeros(sum(P[i].numel() for i in ln_idx), dtype=torch.float32, device=dev)
    off = 0
    for i in ln_idx:
        GP[i] = lnbuf[off:off + P[i].numel()].view_as(P[i])
        off += P[i].numel()
    for l in bs["lins"]:
        GP[l["w"]], GP[l["b"]] = torch.empty_like(P[l["w"]]), torch.empty_like(P[l["b"]])
    _pack_batch(bs["lins"], P, None, G, GP, backward=True)
    for i, pi in enumerate(bs["tables"]):
        GP[pi] = G[-1 - i]
    return GP


class RSTBFunction(torch.autograd.Function):
    """One RSTB of the vanilla SwinIR (swin_transformer_sr.py:471-472): depth Swin blocks at C = 60 -> 3x3 conv -> + input."""

    @staticmethod
    def forward(ctx, bs, geom, X, conv_w, conv_b, *P):
        B, H, Wd = geom
        T = B * H * Wd
        e, z = _f32(X.device)
        _MODE.tc = tc = bs["tc"]
        with _on(X.device):
            W = _link_packed(bs, P, conv_w, conv_b)
            src, saved = X.contiguous(), []
            for st in bs["stl"]:
                sv = _stl_forward(st, W, src, B, H, Wd, tc)
                saved.append(sv)
                src = sv["y"]
            Xn = e(T, 64)
            conv(src, W[bs["lff_w"]], W[bs["lff_b"]], Xn, B, H, Wd, 64, 64, resid=saved[0]["x"])
        ctx.bs, ctx.geom, ctx.W, ctx.P, ctx.saved = bs, geom, W, P, saved
        return Xn

    @staticmethod
    def backward(ctx, dXn):
        bs, W, P, saved = ctx.bs, ctx.W, ctx.P, ctx.saved
        B, H, Wd = ctx.geom
        T = B * H * Wd
        e, z = _f32(dXn.device)
        _MODE.tc = bs["tc"]
        with _on(dXn.device):
            G = _link_grad_buffers(bs, W, P)
            gz = lambda i: G[i]
            dX = dXn.contiguous()
            gemm_tn(dX, saved[-1]["y"], gz(bs["lff_w"]), gz(bs["lff_b"]), 64, 9 * 64, (B, H, Wd, 64))
            dY = e(T, 64)
            conv(dX, conv_dgrad_weight(W[bs["lff_w"]]), z(64), dY, B, H, Wd, 64, 64)
            for k in range(len(bs["stl"]) - 1, -1, -1):
                dY = _stl_backward(bs["stl"][k], saved[k], W, gz, dY, B, H, Wd)
            axpy(dX, dY, 64)                                     # the RSTB residual
            GP = _link_param_grads(bs, P, G)
        ctx.saved = None
        return (None, None, dY, G[bs["lff_w"]], G[bs["lff_b"]]) + tuple(GP)


class SwinIRTailFunction(torch.autograd.Function):
    """norm -> conv_after_body + conv_first skip -> UpsampleOneStep conv (60 -> s^2) (swin_transformer_sr.py:781-799).
    Returns the [T][16] map of sub-pixel channels; PixelShuffle / img_range / mean are views and scalars outside."""

    @staticmethod
    def forward(ctx, tc, geom, X, F0, norm_g, norm_b, cab_w, cab_b, up_w, up_b):
        B, H, Wd = geom
        T = B * H * Wd
        e, z = _f32(X.device)
        _MODE.tc = tc
        with _on(X.device):
            X = X.contiguous()
            FN = z(T, 64)
            _call("rdst_layernorm_fwd", _p(X), 64, _p(norm_g), _p(norm_b), _p(FN), 64, T, 60, 1.0, F32, _lib.stream_ptr())
            F1 = e(T, 64)
            conv(FN, cab_w, cab_b, F1, B, H, Wd, 64, 64, resid=F0)
            up = e(T, 16)
            conv(F1, up_w, up_b, up, B, H, Wd, 64, 16)
        ctx.tc, ctx.geom, ctx.saved = tc, geom, (X, FN, F1, norm_g, cab_w, up_w)
        return up

    @staticmethod
    def backward(ctx, dup):
        B, H, Wd = ctx.geom
        T = B * H * Wd
        X, FN, F1, norm_g, cab_w, up_w = ctx.saved
        e, z = _f32(dup.device)
        _MODE.tc = ctx.tc
        with _on(dup.device):
            dup = dup.contiguous()
            g_upw, g_upb = z(16, 9, 64), z(16)
            gemm_tn(dup, F1, g_upw, g_upb, 16, 9 * 64, (B, H, Wd, 64))
            dF1 = e(T, 64)
            conv(dup, conv_dgrad_weight(up_w), z(64), dF1, B, H, Wd, 16, 64)
            g_cabw, g_cabb = z(64, 9, 64), z(64)
            gemm_tn(dF1, FN, g_cabw, g_cabb, 64, 9 * 64, (B, H, Wd, 64))
            dFN = e(T, 64)
            conv(dF1, conv_dgrad_weight(cab_w), z(64), dFN, B, H, Wd, 64, 64)
            dX, gg, gb = z(T, 64), z(60), z(60)
            _call("rdst_layernorm_bwd", _p(dFN), 64, _p(X), 64, _p(norm_g), _p(dX), 64, _p(gg), _p(gb), T, 60, 1.0,
                  _lib.stream_ptr())
        ctx.saved = None
        return None, None, dX, dF1, gg, gb, g_cabw, g_cabb, g_upw, g_upb


def _check_trainable(m):
    """The training kernels read the parameters through raw pointers as contiguous fp32: anything else (a model cast with
    .half() / .bfloat16() / .double(), a non-contiguous view) must fail loudly instead of being reinterpreted."""
    for name, p in m.named_parameters():
        if p.dtype != torch.float32 or not p.is_contiguous():
            raise TypeError(f"rdst_b200: training needs contiguous float32 parameters; {name} is {p.dtype}"
                            f"{'' if p.is_contiguous() else ', non-contiguous'} (keep the module in fp32: precision='bf16' "
                            "selects bf16 tensor-core GEMMs with fp32 master weights)")


def forward_with_grad_swinir(executor, x):
    """Training forward of rdst_b200.SwinIR: head -> RSTB chain -> tail, same machinery as forward_with_grad."""
    m = executor._module()
    _check_trainable(m)
    if m.training and float(m.drop_path_rate) > 0:
        # the reference applies per-sample stochastic depth (DropPath, swin_transformer_sr.py:199,271-272,686) in train()
        # mode; silently training without it would be a different regularisation
        raise NotImplementedError("rdst_b200.SwinIR: stochastic depth (drop_path_rate > 0) is not implemented for training; "
                                  "construct with drop_path_rate=0 (ini: sir_drop_path_rate = 0.) or call .eval()")
    tc = m.precision == "bf16"
    if tc and not _lib.load().rdst_has_tcgen05():
        raise RuntimeError("rdst_b200: precision='bf16' training needs the tcgen05 kernels (sm_100a device)")
    if x.requires_grad:
        raise NotImplementedError("rdst_b200: gradients with respect to the input image are not implemented")
    dev = x.device
    B, _, H, Wd = x.shape
    geom = (B, H, Wd)
    rng, mean, s = float(m.img_range), float(m.mean.reshape(-1)[0]), m.upscale
    sc = dict(in_scale=rng, in_bias=-mean * rng, tc=tc)
    with packing.differentiable():
        f = packing._f
        hw = torch.zeros(64, 9, 16, device=dev)
        hw[:60, :, 0] = f(m.conv_first.weight).reshape(60, 9)
        hb = torch.zeros(64, device=dev)
        hb[:60] = f(m.conv_first.bias)
        head = [hw, hb, f(m.patch_embed.norm.weight).contiguous(), f(m.patch_embed.norm.bias).contiguous()]
    F0, X = HeadFunction.apply(dict(head_w=0, head_b=1, pe_g=2, pe_b=3), sc, x.detach().to(torch.float32).contiguous(), *head)
    for layer in m.layers:
        bs = dict(rstb_layout(layer), tc=tc)
        cw, cb = pack_conv64(layer.conv, dev)
        X = RSTBFunction.apply(bs, geom, X, cw, cb, *rstb_params(layer))
    cabw, cabb = pack_conv64(m.conv_after_body, dev)
    upw, upb = pack_conv64(m.upsample[0], dev, 16)
    with packing.differentiable():
        ng, nb = packing._f(m.norm.weight).contiguous(), packing._f(m.norm.bias).contiguous()
    up = SwinIRTailFunction.apply(tc, geom, X, F0, ng, nb, cabw, cabb, upw, upb)
    out = up[:, :s * s].reshape(B, H, Wd, s, s).permute(0, 1, 3, 2, 4).reshape(B, 1, H * s, Wd * s)
    return out / rng + mean


class ConvResidualFunction(torch.autograd.Function):
    """Closing step of an RRDSTB (rdst_variations.py:553-555):  out = scale * conv3x3(X) + S  on [T][64] maps."""

    @staticmethod
    def forward(ctx, tc, scale, geom, X, S, w, b):
        e, _ = _f32(X.device)
        _MODE.tc = tc
        with _on(X.device):
            X = X.contiguous()
            out = e(X.shape[0], 64)
            conv(X, w, b, out, *geom, 64, 64, scale=scale, resid=S.contiguous())
        ctx.tc, ctx.scale, ctx.geom, ctx.saved = tc, scale, geom, (X, w)
        return out

    @staticmethod
    def backward(ctx, dout):
        X, w = ctx.saved
        e, z = _f32(dout.device)
        _MODE.tc = ctx.tc
        with _on(dout.device):
            dout = dout.contiguous()
            dy = dout if ctx.scale == 1.0 else dout * ctx.scale
            gw, gb = torch.zeros_like(w), z(64)
            gemm_tn(dy, X, gw, gb, 64, 9 * 64, (*ctx.geom, 64))
            dX = e(X.shape[0], 64)
            conv(dy, conv_dgrad_weight(w), z(64), dX, *ctx.geom, 64, 64)
        ctx.saved = None
        return None, None, None, dX, dout, gw, gb


class BottleneckFunction(torch.autograd.Function):
    """Global bottleneck 'mlp' of RDSTSR_N (rdst_variations.py:1071-1079, 1092-1093) on the [T][64n] concatenation of the RDSTB
    outputs:  F1 = grs * Linear2(Linear1(cat)) + F0."""

    @staticmethod
    def forward(ctx, tc, grs, geom, cat, F0, w1, b1, w2, b2):
        """geom = (B, H, W) selects the 'conv' mode (w2 is a packed 3x3 filter [64][9][64]); None = 'mlp' (w2 [64][64])."""
        T, K = cat.shape
        e, _ = _f32(cat.device)
        _MODE.tc = tc
        with _on(cat.device):
            cat = cat.contiguous()
            tmp, F1 = e(T, 64), e(T, 64)
            linear(cat, w1, b1, tmp, K, 64)
            if geom is None:
                linear(tmp, w2, b2, F1, 64, 64, scale=grs, resid=F0)
            else:
                conv(tmp, w2, b2, F1, *geom, 64, 64, scale=grs, resid=F0)
        ctx.tc, ctx.grs, ctx.geom, ctx.saved = tc, grs, geom, (cat, tmp, w1, w2)
        return F1

    @staticmethod
    def backward(ctx, dF1):
        cat, tmp, w1, w2 = ctx.saved
        T, K = cat.shape
        e, z = _f32(dF1.device)
        _MODE.tc = ctx.tc
        with _on(dF1.device):
            dF1 = dF1.contiguous()
            dy = dF1 if ctx.grs == 1.0 else dF1 * ctx.grs
            gw2, gb2, gw1, gb1 = torch.zeros_like(w2), z(64), z(64, K), z(64)
            dtmp = e(T, 64)
            if ctx.geom is None:
                gemm_tn(dy, tmp, gw2, gb2, 64, 64)
                linear_t(dy, w2, dtmp, 64, 64)
            else:
                gemm_tn(dy, tmp, gw2, gb2, 64, 9 * 64, (*ctx.geom, 64))
                conv(dy, conv_dgrad_weight(w2), z(64), dtmp, *ctx.geom, 64, 64)
            gemm_tn(dtmp, cat, gw1, gb1, 64, K)
            dcat = e(T, K)
            linear_t(dtmp, w1, dcat, 64, K)
        ctx.saved = None
        return None, None, None, dcat, dF1, gw1, gb1, gw2, gb2


def forward_with_grad(executor, x):
    m = executor._module()
    _check_trainable(m)
    if getattr(m, "resi_connection", "1conv") != "1conv" or getattr(m, "dim_modify_mode", "tail") != "tail":
        raise NotImplementedError("rdst_b200: training with resi_connection='3conv' or dim_modify_mode='head' is not implemented "
                                  "(inference only); the E1 configuration trains with '1conv' / 'tail'")
    tc = m.precision == "bf16"
    if tc and not _lib.load().rdst_has_tcgen05():
        raise RuntimeError("rdst_b200: precision='bf16' training needs the tcgen05 kernels (sm_100a device)")
    if x.requires_grad:
        raise NotImplementedError("rdst_b200: gradients with respect to the input image are not implemented")
    dev = x.device
    B, _, H, Wd = x.shape
    geom = (B, H, Wd)
    sc = dict(frozen_scalars(m), tc=tc)
    flat, spec = pack_head(m, dev)
    F0, X = HeadFunction.apply(spec, sc, x.detach().to(torch.float32).contiguous(), *flat)
    feats = []
    nested = hasattr(m.body[0], "residual_scale")            # ESTSR: body.i is an RRDSTB = RDSTBs + conv + shortcut
    for grp in m.body:
        short = X
        for blk in (grp.body if nested else [grp]):
            bs = dict(block_layout(m, blk), tc=tc)
            lff_w, lff_b = pack_lff(blk, dev)
            X = BlockFunction.apply(bs, geom, X, lff_w, lff_b, *block_params(blk))
            feats.append(X)
        if nested:
            cw, cb = pack_conv64(grp.conv, dev)
            X = ConvResidualFunction.apply(tc, float(grp.residual_scale), geom, X, short, cw, cb)
    flat, spec = pack_tail(m, dev)
    if getattr(m, "do_global_bottleneck", False):            # RDSTSR_N: cat -> two Linears; norm / conv_after_body unused
        n = len(feats)
        with packing.differentiable():
            f = packing._f
            pos = torch.cat([torch.arange(60, device=dev) + 64 * i for i in range(n)])
            w1 = torch.zeros(64, 64 * n, device=dev)
            w1[:60, pos] = f(m.bottleneck[0].weight).reshape(60, 60 * n)          # Linear weight or 1x1 conv filter
            b1 = torch.zeros(64, device=dev)
            b1[:60] = f(m.bottleneck[0].bias)
            if m.global_bottleneck_mode == "conv":
                w2, b2 = packing.pack_conv(m.bottleneck[1].weight, m.bottleneck[1].bias, torch.arange(60, device=dev), 64, 64)
            else:
                w2 = torch.zeros(64, 64, device=dev)
                w2[:60, :60] = f(m.bottleneck[1].weight)
                b2 = torch.zeros(64, device=dev)
                b2[:60] = f(m.bottleneck[1].bias)
        F1 = BottleneckFunction.apply(tc, float(m.global_res_scale), geom if m.global_bottleneck_mode == "conv" else None,
                                      torch.cat(feats, 1), F0, w1, b1, w2, b2)
        return TailFunction.apply(spec, dict(sc, tail_only=True), geom, F1, F0, *flat)
    return TailFunction.apply(spec, sc, geom, X, F0, *flat)
