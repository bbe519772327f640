"""CPU stand-ins for the librdst_b200 entry points.  TEST INFRASTRUCTURE ONLY (used by `-m "not gpu"` tests).

Each function restates the *contract* written in include/rdst_b200.h with plain torch ops on CPU tensors, so the
host-side logic (weight packing, padded layouts, launch order, buffer reuse) can be checked against the golden
vectors in a container without a GPU.  It is never importable from the product package: tests monkeypatch
`rdst_b200._lib.call/ptr/stream_ptr` for the duration of one test.  The real kernels are checked against the
oracle on the GPU by the `-m gpu` tests.
"""
import contextlib

import torch
import torch.nn.functional as F


def _store(y, val):
    y[:, :val.shape[1]] = val.to(y.dtype)


def _lnhat(x, creal):
    k = x.shape[1]
    mean = x.sum(1, keepdim=True) / creal
    ss = ((x - mean) ** 2).sum(1, keepdim=True) - (k - creal) * mean * mean
    return (x - mean) * torch.rsqrt(ss.clamp_min(0) / creal + 1e-5)


def rdst_linear_fwd(x, ldx, w, bias, resid, ldr, y, ldy, T, K, N, ln_creal, act, out_scale, dt, st):
    assert x.stride(0) == ldx and y.stride(0) == ldy and x.shape[0] == T
    a = x[:, :K].float()
    if ln_creal > 0:
        a = _lnhat(a, ln_creal)
    assert w.shape == (N, K), (w.shape, N, K)
    v = a @ w.t() + bias
    if act == 1:
        v = F.gelu(v)
    elif act == 2:
        v = F.leaky_relu(v, 0.2)
    v = v * out_scale
    if resid is not None:
        assert resid.stride(0) == ldr
        v = v + resid[:, :N].float()
    _store(y, v)


def rdst_window_attention_fwd(qkv, ldq, table, out, ldo, B, H, W, C, heads, shift, dt, st):
    hd = C // heads
    t = qkv[:, :3 * C].float().reshape(B, H, W, 3 * C)
    if shift:
        t = torch.roll(t, (-shift, -shift), (1, 2))
    win = t.reshape(B, H // 8, 8, W // 8, 8, 3 * C).permute(0, 1, 3, 2, 4, 5).reshape(-1, 64, 3, heads, hd)
    q, k, v = (win[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    att = q @ k.transpose(-1, -2)
    r = torch.arange(8)
    ih, iw = (g.reshape(-1) for g in torch.meshgrid(r, r, indexing="ij"))
    idx = (ih[:, None] - ih[None, :] + 7) * 15 + (iw[:, None] - iw[None, :] + 7)
    att = att + table[idx.reshape(-1)].reshape(64, 64, heads).permute(2, 0, 1)[None]
    if shift:
        hs = torch.arange(H)[:, None].expand(H, W)
        ws = torch.arange(W)[None, :].expand(H, W)
        reg = (torch.where(hs < H - 8, 0, torch.where(hs < H - shift, 1, 2)) * 3 +
               torch.where(ws < W - 8, 0, torch.where(ws < W - shift, 1, 2)))
        rw = reg.reshape(H // 8, 8, W // 8, 8).permute(0, 2, 1, 3).reshape(-1, 64)
        mask = torch.where(rw[:, :, None] != rw[:, None, :], -100.0, 0.0)          # (nW,64,64)
        att = att + mask.repeat(B, 1, 1)[:, None]
    o = (torch.softmax(att, -1) @ v).permute(0, 2, 1, 3).reshape(-1, 64, C)
    o = o.reshape(B, H // 8, W // 8, 8, 8, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    _store(out, o.reshape(-1, C))


def rdst_conv3x3_act_fwd(x, ldx, w, bias, y, ldy, B, H, W, Cin, N, act, dt, st):
    a = x[:, :Cin].float().reshape(B, H, W, Cin).permute(0, 3, 1, 2)
    assert w.shape == (N, 9, Cin)
    v = F.conv2d(a, w.reshape(N, 3, 3, Cin).permute(0, 3, 1, 2), bias, padding=1)
    v = F.gelu(v) if act == 1 else (F.leaky_relu(v, 0.2) if act == 2 else v)
    _store(y, v.permute(0, 2, 3, 1).reshape(-1, N))


def rdst_conv3x3_fwd(x, ldx, w, bias, resid, ldr, y, ldy, B, H, W, Cin, N, out_scale, shuffle, dt, st):
    a = x[:, :Cin].float().reshape(B, H, W, Cin).permute(0, 3, 1, 2)
    assert w.shape == (N, 9, Cin)
    wt = w.reshape(N, 3, 3, Cin).permute(0, 3, 1, 2)
    v = F.conv2d(a, wt, bias, padding=1) * out_scale           # (B,N,H,W)
    if shuffle == 2:
        G = N // 4
        v = v.reshape(B, 2, 2, G, H, W).permute(0, 4, 1, 5, 2, 3).reshape(B * 2 * H * 2 * W, G)
        _store(y, v)
    else:
        v = v.permute(0, 2, 3, 1).reshape(-1, N)
        if resid is not None:
            v = v + resid[:, :N].float()
        _store(y, v)


def rdst_head_fwd(img, in_scale, in_bias, w, bias, gamma, beta, feat0, ldf, dense, ldd, B, H, W, dt, st):
    a = img.float() * in_scale + in_bias
    f = F.conv2d(a, w.reshape(60, 1, 3, 3), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, 60)
    z = torch.zeros(f.shape[0], 4)
    _store(feat0, torch.cat([f, z], 1))
    _store(dense, torch.cat([F.layer_norm(f, (60,), gamma, beta, 1e-5), z], 1))


def rdst_layernorm_fwd(x, ldx, gamma, beta, y, ldy, T, creal, out_scale, dt, st):
    _store(y, F.layer_norm(x[:, :creal].float(), (creal,), gamma, beta, 1e-5) * out_scale)


def rdst_last_conv_fwd(x, ldx, w, bias, out_scale, out_bias, img, B, H, W, Cin, dt, st):
    a = x[:, :Cin].float().reshape(B, H, W, Cin).permute(0, 3, 1, 2)
    wt = w.reshape(1, 3, 3, Cin).permute(0, 3, 1, 2)
    img.copy_((F.conv2d(a, wt, None, padding=1) + bias) * out_scale + out_bias)


# ---------------------------------------------------------------------------------------------------------------
# training-path entry points (fp32 contract; the tensor-core variants compute the same thing on bf16-rounded operands)
# ---------------------------------------------------------------------------------------------------------------
def _dense_real(k):
    idx = torch.arange(k)
    return (idx < 60) | ((idx >= 64) & (((idx - 64) % 32) < 30))


def _real_mask(K, creal, dense_layout):
    return _dense_real(K) if dense_layout else (torch.arange(K) < creal)


def rdst_gelu_fwd(x, ldx, y, ldy, T, N, st):
    _store(y, F.gelu(x[:, :N].float()))


@torch.enable_grad()          # called from inside autograd.Function.backward, where grad mode is off
def rdst_gelu_bwd(x, ldx, dy, ldd, dx, ldo, T, N, st):
    xv = x[:, :N].double().clone().requires_grad_(True)
    g, = torch.autograd.grad(F.gelu(xv), xv, dy[:, :N].double())
    _store(dx, g)


def _conv_cols(x, B, H, W, Cin):
    xi = x[:, :Cin].double().reshape(B, H, W, Cin).permute(0, 3, 1, 2)
    return F.unfold(xi, 3, padding=1).reshape(B, Cin, 9, H * W).permute(0, 3, 2, 1).reshape(B * H * W, 9 * Cin)   # [t][tap][ci]


def rdst_gemm_tn_acc(dy, ldy, x, ldx, dw, db, T, N, K, conv, B, H, W, Cin, st):
    xs = _conv_cols(x, B, H, W, Cin) if conv else x[:, :K].double()
    g = dy[:, :N].double().t() @ xs
    dw.view(N, K).add_(g.to(dw.dtype))
    if db is not None:
        db[:N].add_(dy[:, :N].double().sum(0).to(db.dtype))


def rdst_lnhat_fwd(x, ldx, y, ldy, T, K, creal, dense_layout, st):
    _store(y, _lnhat(x[:, :K].float(), creal) * _real_mask(K, creal, dense_layout))


@torch.enable_grad()          # called from inside autograd.Function.backward, where grad mode is off
def rdst_lnhat_bwd(dxh, ldd, x, ldx, resid, ldr, resid2, ldr2, dx, ldo, T, K, creal, dense_layout, st):
    real = _real_mask(K, creal, dense_layout).double()
    xv = x[:, :K].double().clone().requires_grad_(True)
    xh = _lnhat(xv, creal) * real
    g, = torch.autograd.grad(xh, xv, dxh[:, :K].double() * real)
    v = g * real                      # pad outputs are 0 (the kernel does not propagate into pad channels)
    if resid is not None:
        v = v + resid[:, :K].double()
    if resid2 is not None:
        v = v + resid2[:, :K].double()
    _store(dx, v)


@torch.enable_grad()          # called from inside autograd.Function.backward, where grad mode is off
def rdst_layernorm_bwd(dy, ldd, x, ldx, gamma, dx, ldo, dgamma, dbeta, T, creal, scale, st):
    xv = x[:, :creal].double().clone().requires_grad_(True)
    gv = gamma.double().clone().requires_grad_(True)
    bv = torch.zeros(creal, dtype=torch.float64, requires_grad=True)
    y = F.layer_norm(xv, (creal,), gv, bv, 1e-5) * scale
    gx, gg, gb = torch.autograd.grad(y, (xv, gv, bv), dy[:, :creal].double())
    _store(dx, gx)
    dgamma.add_(gg.to(dgamma.dtype))
    dbeta.add_(gb.to(dbeta.dtype))


@torch.enable_grad()          # called from inside autograd.Function.backward, where grad mode is off
def rdst_window_attention_bwd(qkv, ldq, table, dout, ldo, dqkv, ldg, dtable, B, H, W, C, heads, shift, st):
    q = qkv[:, :3 * C].double().clone().requires_grad_(True)
    t = table.double().clone().requires_grad_(True)
    o = torch.zeros(q.shape[0], C, dtype=torch.float64)
    o = _attention_core(q, t, B, H, W, C, heads, shift)
    gq, gt = torch.autograd.grad(o, (q, t), dout[:, :C].double())
    _store(dqkv, gq)
    dtable.add_(gt.to(dtable.dtype))


def _attention_core(qkv, table, B, H, W, C, heads, shift):
    hd = C // heads
    t = qkv.reshape(B, H, W, 3 * C)
    if shift:
        t = torch.roll(t, (-shift, -shift), (1, 2))
    win = t.reshape(B, H // 8, 8, W // 8, 8, 3 * C).permute(0, 1, 3, 2, 4, 5).reshape(-1, 64, 3, heads, hd)
    q, k, v = (win[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    att = q @ k.transpose(-1, -2)
    r = torch.arange(8)
    ih, iw = (g.reshape(-1) for g in torch.meshgrid(r, r, indexing="ij"))
    idx = (ih[:, None] - ih[None, :] + 7) * 15 + (iw[:, None] - iw[None, :] + 7)
    att = att + table[idx.reshape(-1)].reshape(64, 64, heads).permute(2, 0, 1)[None]
    if shift:
        hs = torch.arange(H)[:, None].expand(H, W)
        ws = torch.arange(W)[None, :].expand(H, W)
        reg = (torch.where(hs < H - 8, 0, torch.where(hs < H - shift, 1, 2)) * 3 +
               torch.where(ws < W - 8, 0, torch.where(ws < W - shift, 1, 2)))
        rw = reg.reshape(H // 8, 8, W // 8, 8).permute(0, 2, 1, 3).reshape(-1, 64)
        mask = torch.where(rw[:, :, None] != rw[:, None, :], -100.0, 0.0).to(att.dtype)
        att = att + mask.repeat(B, 1, 1)[:, None]
    o = (torch.softmax(att, -1) @ v).permute(0, 2, 1, 3).reshape(-1, 64, C)
    o = o.reshape(B, H // 8, W // 8, 8, 8, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    return o.reshape(-1, C)


def rdst_axpy(x, ldx, y, ldy, T, N, alpha, st):
    y[:, :N].add_(x[:, :N], alpha=alpha)


def rdst_pixel_unshuffle2(u, ldu, z, ldz, B, H, W, G, st):
    v = u[:, :G].reshape(B, H, 2, W, 2, G).permute(0, 1, 3, 2, 4, 5).reshape(B * H * W, 4 * G)     # [t][s = 2dy+dx][c]
    _store(z, v)


def _chan_pos(n, scatter):
    idx = torch.arange(n)
    if not scatter:
        return idx
    g = torch.clamp((idx - 60) // 30, min=0)
    return torch.where(idx < 60, idx, 64 + 32 * g + (idx - 60) % 30)


def rdst_pack_linear_batch(descs, n, backward, st):
    """descs: list of dicts with TENSORS under the RdstPackDesc field names (the emulator replaces _lib.pack_desc_array)."""
    for d in descs[:n]:
        N, K = d["N"], d["K"]
        W = d["W"].double().reshape(N, K)
        rs = torch.ones(N, dtype=torch.float64)
        rs[:d["q_rows"]] = d["q_scale"]
        pn, pk = _chan_pos(N, d["scatter_rows"]), _chan_pos(K, d["scatter_cols"])
        g = d["gamma"].double() if d["gamma"] is not None else torch.ones(K, dtype=torch.float64)
        be = d["beta"].double() if d["beta"] is not None else torch.zeros(K, dtype=torch.float64)
        if not backward:
            Wp = d["Wp"]
            Wp[pn[:, None], pk[None, :]] = (rs[:, None] * W * g[None, :]).to(Wp.dtype)
            d["bp"][pn] = (rs * ((d["b"].double() if d["b"] is not None else 0) + W @ be)).to(d["bp"].dtype)
        else:
            gw = d["dWp"].double()[pn[:, None], pk[None, :]]
            gb = d["dbp"].double()[pn]
            d["dW"].copy_((rs[:, None] * (g[None, :] * gw + gb[:, None] * be[None, :])).reshape(d["dW"].shape))
            if d["db"] is not None:
                d["db"].copy_(rs * gb)
            if d["dgamma"] is not None:
                d["dgamma"].add_(((rs[:, None] * W * gw).sum(0)).to(d["dgamma"].dtype))
                d["dbeta"].add_(((rs[:, None] * W) * gb[:, None]).sum(0).to(d["dbeta"].dtype))


# ---- tensor-core entry points of the bf16 training mode: same contracts, evaluated here WITHOUT the bf16 operand rounding
#      (the rounding is a property of the kernels, checked on the GPU; the host logic -- padded strides, fused operands,
#      saved log-sum-exp, descriptor plumbing -- is what these stand-ins let the CPU tests exercise)
def _gelu_grad(x):
    return 0.5 * (1 + torch.erf(x * 0.7071067811865476)) + x * torch.exp(-0.5 * x * x) * 0.3989422804014327


def rdst_gemm_tc(x, ldx, w, ldw, w_mn, bias, resid, ldr, aux, lda, y, ldy, T, K, N, a_op, ln_creal, scale, conv, B, H, W, Cin,
                 shuffle, st):
    assert x.stride(0) == ldx and y.stride(0) == ldy and x.data_ptr() % 16 == 0 and y.data_ptr() % 16 == 0
    assert ldx % 4 == 0 and ldy % 4 == 0 and (conv or ldx >= (K + 7) // 8 * 8)
    if conv:
        assert not w_mn and not a_op
        a = _conv_cols(x, B, H, W, Cin)
        wm = w.double().reshape(-1, 9 * Cin)[:N]
    else:
        a = x[:, :K].double()
        if a_op == 1:
            a = _lnhat(a, ln_creal)
        elif a_op == 2:
            a = F.gelu(a)
        wm = (w.double()[:K, :N].t() if w_mn else w.double().reshape(-1, w.shape[-1])[:N, :K])
    v = a @ wm.t()
    if bias is not None:
        v = v + bias.double()[:N]
    v = v * scale
    if aux is not None:
        v = v * _gelu_grad(aux[:, :N].double())
    if shuffle == 2:
        G = N // 4
        v = v.reshape(B, H, W, 2, 2, G).permute(0, 1, 3, 2, 4, 5).reshape(B * 2 * H * 2 * W, G)
    elif resid is not None:
        v = v + resid[:, :N].double()
    _store(y, v)


def rdst_gemm_tc_lnbwd(dy, ldy, w, ldw, x, ldx, resid, ldr, resid2, ldr2, dx, ldo, T, K, N, creal, scale, st):
    assert N <= 128 and N % 16 == 0 and dy.stride(0) == ldy and dx.stride(0) == ldo
    dxh = (dy[:, :K].double() @ w.double()[:K, :N]) * scale
    rdst_lnhat_bwd(dxh, N, x, ldx, resid, ldr, resid2, ldr2, dx, ldo, T, N, creal, 1, st)


def rdst_gemm_tn_tc(dy, ldy, x, ldx, dw, db, T, N, K, conv, B, H, W, Cin, x_op, x_creal, st):
    assert dy.stride(0) == ldy and ldy % 4 == 0 and ldy >= (N + 7) // 8 * 8 and dw.data_ptr() % 16 == 0
    if x_op:
        xs = x[:, :K].double()
        x = (_lnhat(xs, x_creal) if x_op == 1 else F.gelu(xs)).contiguous()
    rdst_gemm_tn_acc(dy, ldy, x, x.stride(0), dw, db, T, N, K, conv, B, H, W, Cin, st)


def rdst_window_attention_tc_fwd(qkv, ldq, table, out, ldo, lse, B, H, W, C, shift, st):
    assert ldq >= 3 * C and ldo >= C
    rdst_window_attention_fwd(qkv, ldq, table, out, ldo, B, H, W, C, 6, shift, 0, st)
    if lse is not None:
        lse.fill_(float("nan"))          # the stand-in backward recomputes the softmax; a real lse must never be needed here


def rdst_window_attention_tc_bwd(qkv, ldq, table, lse, dout, ldo, dqkv, ldg, dtable, B, H, W, C, shift, st):
    assert lse is not None and lse.shape == (qkv.shape[0], 6)
    rdst_window_attention_bwd(qkv, ldq, table, dout, ldo, dqkv, ldg, dtable, B, H, W, C, 6, shift, st)


_TABLE = {k: v for k, v in globals().items() if k.startswith("rdst_")}


@contextlib.contextmanager
def emulated_abi(record=None):
    """Route rdst_b200._lib.call to the CPU stand-ins above; pointers become the tensors themselves."""
    from rdst_b200 import _lib

    def call(name, *args):
        if record is not None:
            record.append(name)
        _TABLE[name](*args)

    import types
    saved = (_lib.call, _lib.ptr, _lib.stream_ptr, _lib.pack_desc_array, _lib.load)
    _lib.call, _lib.ptr, _lib.stream_ptr, _lib.pack_desc_array = call, (lambda t: t), (lambda: None), (lambda descs: descs)
    _lib.load = lambda: types.SimpleNamespace(rdst_has_tcgen05=lambda: 1)     # capability probe of the bf16 training mode
    try:
        yield
    finally:
        _lib.call, _lib.ptr, _lib.stream_ptr, _lib.pack_desc_array, _lib.load = saved
