"""Training path: forward with saved activations + hand-written backward over librdst_b200 kernels (fp32).

The whole network is ONE torch.autograd.Function.  Its tensor inputs are the *packed* weights
(rdst_b200/packing.py run in differentiable mode: LayerNorm folding, q scaling, padded scatter are ordinary torch ops
on the 4.5 M parameters), so autograd carries the gradients of the packed tensors back to the reference-named
parameters; every activation-sized computation, forward and backward, is a kernel of the C ABI:
    data gradients     rdst_linear_fwd / rdst_conv3x3_fwd with transposed (and tap-flipped) weights
    weight gradients   rdst_gemm_tn_acc (token reduction, optional 3x3 gather) incl. bias column sums
    LayerNorm          rdst_lnhat_fwd/_bwd (affine folded), rdst_layernorm_bwd (patch_embed.norm, final norm)
    attention          rdst_window_attention_bwd (recomputes the probabilities; relative-position table gradient)
    GELU, residual joins, PixelShuffle  rdst_gelu_fwd/_bwd, rdst_axpy, rdst_pixel_unshuffle2
Gradient of: RDSTSR.forward (rdst_variations.py:1342-1360) and everything it calls; checked against torch.autograd
through the CPU oracle in tests/test_gpu_backward.py.  Training runs in precision='fp32' (this round).
"""
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
def linear(x, w, b, y, K, N, ln_creal=0, act=0, scale=1.0, resid=None):
    T = x.shape[0]
    _call("rdst_linear_fwd", _p(x), _ld(x), _p(w), _p(b), _p(resid), 0 if resid is None else _ld(resid), _p(y), _ld(y),
          T, K, N, ln_creal, act, scale, F32, _lib.stream_ptr())


def conv(x, w, b, y, B, H, W, cin, n, scale=1.0, shuffle=0, resid=None):
    _call("rdst_conv3x3_fwd", _p(x), _ld(x), _p(w), _p(b), _p(resid), 0 if resid is None else _ld(resid), _p(y), _ld(y),
          B, H, W, cin, n, scale, shuffle, F32, _lib.stream_ptr())


def gemm_tn(dy, x, dw, db, N, K, conv_geom=None):
    T = dy.shape[0]
    if conv_geom is None:
        _call("rdst_gemm_tn_acc", _p(dy), _ld(dy), _p(x), _ld(x), _p(dw), _p(db), T, N, K, 0, 0, 0, 0, 0, _lib.stream_ptr())
    else:
        B, H, W, cin = conv_geom
        _call("rdst_gemm_tn_acc", _p(dy), _ld(dy), _p(x), _ld(x), _p(dw), _p(db), T, N, K, 1, B, H, W, cin, _lib.stream_ptr())


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
def train_weights(m, device):
    """Differentiable packed weights as a flat list + a spec to rebuild the nested structure."""
    with packing.differentiable():
        flat, spec = [], {"blocks": []}

        def add(t):
            flat.append(t.contiguous())
            return len(flat) - 1

        for blk in m.body:
            bs = {"dstl": []}
            c = packing.EMBED
            for dstl in blk.body:
                stls = []
                for b in dstl.body.blocks:
                    p = packing.pack_stl(b, c)
                    stls.append({k: add(p[k]) for k in ("wqkv", "bqkv", "wproj", "bproj", "w1", "b1", "w2", "b2", "table")}
                                | {"c": c, "cp": p["cp"], "hp": p["hp"], "shift": b.shift_size})
                t = packing.pack_dstl_tail(dstl, c, m.dense_scale)
                bs["dstl"].append({"c": c, "stl": stls, "tw": add(t["w"]), "tb": add(t["b"]), "scale": t["scale"]})
                c += packing.GROWTH
            pos = packing.channel_positions(c, device)
            w, b = packing.pack_conv(blk.conv.weight, blk.conv.bias, pos, packing.DENSE_LD, 64)
            bs["lff_w"], bs["lff_b"] = add(w), add(b)
            spec["blocks"].append(bs)
        id60 = torch.arange(60, device=device)
        f = packing._f
        spec["head_w"] = add(f(m.head.weight).reshape(60, 9))
        spec["head_b"] = add(f(m.head.bias))
        spec["pe_g"], spec["pe_b"] = add(f(m.patch_embed.norm.weight)), add(f(m.patch_embed.norm.bias))
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
    spec["scalars"] = dict(in_scale=float(m.sub_mean.weight.detach().reshape(-1)[0]),
                           in_bias=float(m.sub_mean.bias.detach().reshape(-1)[0]),
                           out_scale=float(m.add_mean.weight.detach().reshape(-1)[0]),
                           out_bias=float(m.add_mean.bias.detach().reshape(-1)[0]),
                           res_scale=float(m.rdb_residual_scale), grs=float(m.global_res_scale),
                           flo=bool(m.feature_last_operation), sr=int(m.sr_scale))
    return flat, spec


# ------------------------------------------------------------------------------------------------ the Function
class RDSTFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec, x, *W):
        dev = x.device
        B, _, H, Wd = x.shape
        T = B * H * Wd
        sc = spec["scalars"]
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        saved = {"blocks": []}
        with torch.cuda.device(dev):
            # head: conv 1->60 on a 16-channel padded image (generic conv kernel), then patch_embed LayerNorm
            img = z(T, 16)
            img[:, 0] = x.reshape(-1) * sc["in_scale"] + sc["in_bias"]
            hw = z(64, 9, 16); hw[:60, :, 0] = W[spec["head_w"]]
            hb = z(64); hb[:60] = W[spec["head_b"]]
            F0 = e(T, 64)
            conv(img, hw, hb, F0, B, H, Wd, 16, 64)
            D = z(T, 160)
            _call("rdst_layernorm_fwd", _p(F0), 64, _p(W[spec["pe_g"]]), _p(W[spec["pe_b"]]), _p(D), 160, T, 60, 1.0, F32,
                  _lib.stream_ptr())
            saved["img"], saved["F0"] = img, F0
            for bs in spec["blocks"]:
                sb = {"D": D, "dstl": []}
                for j, ds in enumerate(bs["dstl"]):
                    c = ds["c"]
                    src = D
                    sl = []
                    for st in ds["stl"]:
                        cp, hp = st["cp"], st["hp"]
                        qkv, o, x1, hid, act, y = e(T, 3 * c), e(T, c), e(T, cp), e(T, hp), e(T, hp), e(T, cp)
                        linear(src, W[st["wqkv"]], W[st["bqkv"]], qkv, cp, 3 * c, ln_creal=c)
                        _call("rdst_window_attention_fwd", _p(qkv), 3 * c, _p(W[st["table"]]), _p(o), c, B, H, Wd, c,
                              packing.HEADS, st["shift"], F32, _lib.stream_ptr())
                        linear(o, W[st["wproj"]], W[st["bproj"]], x1, c, cp, resid=src)
                        linear(x1, W[st["w1"]], W[st["b1"]], hid, cp, hp, ln_creal=c)
                        _call("rdst_gelu_fwd", _p(hid), hp, _p(act), hp, T, hp, _lib.stream_ptr())
                        linear(act, W[st["w2"]], W[st["b2"]], y, hp, cp, resid=x1)
                        sl.append(dict(x=src, qkv=qkv, o=o, x1=x1, hid=hid, y=y))
                        del act
                        src = y
                    off = 64 + 32 * j
                    linear(src, W[ds["tw"]], W[ds["tb"]], D[:, off:], ds["stl"][0]["cp"], 32, ln_creal=c, scale=ds["scale"])
                    sb["dstl"].append(sl)
                Dn = z(T, 160)
                conv(D, W[bs["lff_w"]], W[bs["lff_b"]], Dn, B, H, Wd, 160, 64, scale=sc["res_scale"], resid=D)
                saved["blocks"].append(sb)
                D = Dn
            saved["Dlast"] = D
            FN = z(T, 64)
            _call("rdst_layernorm_fwd", _p(D), 160, _p(W[spec["norm_g"]]), _p(W[spec["norm_b"]]), _p(FN), 64, T, 60,
                  sc["grs"], F32, _lib.stream_ptr())
            F1 = e(T, 64)
            if sc["flo"]:
                conv(FN, W[spec["cab_w"]], W[spec["cab_b"]], F1, B, H, Wd, 64, 64, resid=F0)
            else:
                F1.copy_(FN + F0)
            saved["FN"] = FN
            feats = [F1]
            h, w_ = H, Wd
            for wi, bi in spec["up"]:
                up = e(B * 4 * h * w_, 64)
                conv(feats[-1], W[wi], W[bi], up, B, h, w_, 64, 256, shuffle=2)
                feats.append(up)
                h, w_ = 2 * h, 2 * w_
            saved["feats"] = feats
            out = e(B * h * w_, 1)
            conv(feats[-1], W[spec["last_w"]], W[spec["last_b"]], out, B, h, w_, 64, 1)
            out = (out * sc["out_scale"] + sc["out_bias"]).reshape(B, 1, h, w_)
        ctx.spec, ctx.saved, ctx.W, ctx.geom = spec, saved, W, (B, H, Wd)
        return out

    @staticmethod
    def backward(ctx, dout):
        spec, S, W = ctx.spec, ctx.saved, ctx.W
        B, H, Wd = ctx.geom
        T = B * H * Wd
        sc = spec["scalars"]
        dev = dout.device
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        G = [None] * len(W)

        def gz(i):
            if G[i] is None:
                G[i] = torch.zeros_like(W[i])
            return G[i]

        with torch.cuda.device(dev):
            feats = S["feats"]
            n_up = len(spec["up"])
            h, w_ = H * (2 ** n_up), Wd * (2 ** n_up)
            # ---- last conv (64 -> 1) + add_mean ----
            dy = (dout.to(torch.float32) * sc["out_scale"]).reshape(-1, 1).contiguous()
            gemm_tn(dy, feats[-1], gz(spec["last_w"]), gz(spec["last_b"]), 1, 9 * 64, (B, h, w_, 64))
            dy16 = z(dy.shape[0], 16); dy16[:, 0] = dy[:, 0]
            wT = z(64, 9, 16); wT[:, :, 0] = conv_dgrad_weight(W[spec["last_w"]])[:, :, 0]
            dfeat = e(dy.shape[0], 64)
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
            dF0 = dF1.clone()
            dFN = e(T, 64)
            if sc["flo"]:
                gemm_tn(dF1, S["FN"], gz(spec["cab_w"]), gz(spec["cab_b"]), 64, 9 * 64, (B, H, Wd, 64))
                conv(dF1, conv_dgrad_weight(W[spec["cab_w"]]), z(64), dFN, B, H, Wd, 64, 64)
            else:
                dFN.copy_(dF1)
            # ---- final norm ----
            dX = z(T, 64)                                 # grad wrt the trunk (block output), pads zero
            _call("rdst_layernorm_bwd", _p(dFN), 64, _p(S["Dlast"]), 160, _p(W[spec["norm_g"]]), _p(dX), 64,
                  _p(gz(spec["norm_g"])), _p(gz(spec["norm_b"])), T, 60, sc["grs"], _lib.stream_ptr())
            # ---- RDSTBs in reverse ----
            for bs, sb in zip(reversed(spec["blocks"]), reversed(S["blocks"])):
                D = sb["D"]
                dys = dX if sc["res_scale"] == 1.0 else dX * sc["res_scale"]
                gemm_tn(dys, D, gz(bs["lff_w"]), gz(bs["lff_b"]), 64, 9 * 160, (B, H, Wd, 160))
                dD = e(T, 160)
                conv(dX, conv_dgrad_weight(W[bs["lff_w"]]), z(160), dD, B, H, Wd, 64, 160, scale=sc["res_scale"])
                axpy(dX, dD, 64)                          # residual: block output = LFF(D) + D[:, :64]
                for j in range(len(bs["dstl"]) - 1, -1, -1):
                    ds, sl = bs["dstl"][j], sb["dstl"][j]
                    c = ds["c"]
                    cp = ds["stl"][0]["cp"]
                    off = 64 + 32 * j
                    dg = dD[:, off:off + 32]
                    y1 = sl[-1]["y"]
                    xh = e(T, cp)
                    lnhat(y1, xh, cp, c)
                    gw, gb = z(32, cp), z(32)
                    gemm_tn(dg, xh, gw, gb, 32, cp)
                    gz(ds["tw"]).add_(gw, alpha=ds["scale"]); gz(ds["tb"]).add_(gb, alpha=ds["scale"])
                    dxh = e(T, cp)
                    linear(dg, W[ds["tw"]].t().contiguous(), z(cp), dxh, 32, cp, scale=ds["scale"])
                    dy_cur = e(T, cp)
                    lnhat_bwd(dxh, y1, dy_cur, cp, c)
                    del xh, dxh
                    for k in range(len(ds["stl"]) - 1, -1, -1):
                        first = k == 0
                        dy_cur = _stl_backward(ds["stl"][k], sl[k], W, gz, dy_cur, B, H, Wd,
                                               accumulate_into=dD if first else None)
                dX = dD[:, :64].contiguous()
                del dD
            # ---- patch_embed norm + head conv ----
            dE = e(T, 64)
            dE.zero_()
            _call("rdst_layernorm_bwd", _p(dX), 64, _p(S["F0"]), 64, _p(W[spec["pe_g"]]), _p(dE), 64,
                  _p(gz(spec["pe_g"])), _p(gz(spec["pe_b"])), T, 60, 1.0, _lib.stream_ptr())
            axpy(dE, dF0, 60)
            ghw, ghb = z(64, 9 * 16), z(64)
            gemm_tn(dF0, S["img"], ghw, ghb, 64, 9 * 16, (B, H, Wd, 16))
            gz(spec["head_w"]).add_(ghw.reshape(64, 9, 16)[:60, :, 0])
            gz(spec["head_b"]).add_(ghb[:60])
        ctx.saved = None
        return (None, None) + tuple(G)


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
    act = e(T, hp)
    _call("rdst_gelu_fwd", _p(hid), hp, _p(act), hp, T, hp, _lib.stream_ptr())
    gemm_tn(dY, act, gz(st["w2"]), gz(st["b2"]), cp, hp)
    dact = act                                                      # reuse the buffer
    linear(dY, W[st["w2"]].t().contiguous(), z(hp), dact, cp, hp)
    dhid = e(T, hp)
    _call("rdst_gelu_bwd", _p(hid), hp, _p(dact), hp, _p(dhid), hp, T, hp, _lib.stream_ptr())
    xh = e(T, cp)
    lnhat(x1, xh, cp, c)
    gemm_tn(dhid, xh, gz(st["w1"]), gz(st["b1"]), hp, cp)
    dxh = e(T, cp)
    linear(dhid, W[st["w1"]].t().contiguous(), z(cp), dxh, hp, cp)
    dX1 = e(T, cp)
    lnhat_bwd(dxh, x1, dX1, cp, c, resid=dY)
    del act, dhid
    # ---- x1 = x + proj(attn(lnhat(x))) ----
    gemm_tn(dX1, o, gz(st["wproj"]), gz(st["bproj"]), cp, c)
    dO = e(T, c)
    linear(dX1, W[st["wproj"]].t().contiguous(), z(c), dO, cp, c)
    dqkv = e(T, 3 * c)
    _call("rdst_window_attention_bwd", _p(qkv), 3 * c, _p(W[st["table"]]), _p(dO), c, _p(dqkv), 3 * c,
          _p(gz(st["table"])), B, H, Wd, c, packing.HEADS, st["shift"], _lib.stream_ptr())
    lnhat(x, xh, cp, c)
    gemm_tn(dqkv, xh, gz(st["wqkv"]), gz(st["bqkv"]), 3 * c, cp)
    linear(dqkv, W[st["wqkv"]].t().contiguous(), z(cp), dxh, 3 * c, cp)
    if accumulate_into is None:
        dX = e(T, cp)
        lnhat_bwd(dxh, x, dX, cp, c, resid=dX1)
        return dX
    lnhat_bwd(dxh, x, accumulate_into, cp, c, resid=dX1, resid2=accumulate_into)
    return None


def forward_with_grad(executor, x):
    m = executor._module()
    if m.precision != "fp32":
        raise NotImplementedError("rdst_b200: training (autograd) is implemented for precision='fp32' in this round; "
                                  "call model.set_precision('fp32') for training, bf16 is inference-only for now")
    if x.requires_grad:
        raise NotImplementedError("rdst_b200: gradients with respect to the input image are not implemented")
    flat, spec = train_weights(m, x.device)
    return RDSTFunction.apply(spec, x.detach().to(torch.float32).contiguous(), *flat)
