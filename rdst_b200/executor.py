"""Launch sequence of the RDST forward pass over librdst_b200 kernels.

One `Executor` is bound to one RDSTSR module.  It owns (a) the packed-weight cache (rebuilt when a parameter's
version counter changes, e.g. after optimizer.step() or load_state_dict) and (b) per-shape workspaces allocated
through torch so the caching allocator / CUDA-graph pools see them.  Kernels are enqueued on torch's current
stream; nothing here synchronises.

Data flow per RDSTB (dense buffer D is [T][160], see include/rdst_b200.h):
    for j in 0..2:  C = 60+30j
        STL(shift 0):  D[:, :Cp] -> Y0          STL(shift 4): Y0 -> Y1
        tail:          LN+Linear(Y1) * dense_scale -> D[:, 64+32j : 96+32j]        (the reference's torch.cat)
    LFF: conv3x3(D, 150->60) * res_scale + D[:, :64]  ->  D'[:, :64]               (ping-pong buffer)
"""
import weakref

import torch

from . import _lib, packing


def call(name, *args):
    _lib.call(name, *args)


def ptr(t):
    return _lib.ptr(t)


def _pack_3conv(seq, cin_pos, cin_width, device):
    """Conv 3x3 (Cin -> mid) + LeakyReLU, conv 1x1 (mid -> mid) + LeakyReLU, conv 3x3 (mid -> 60): the '3conv' fusion of
    RDSTB / conv_after_body (rdst_variations.py:422-427, :1286-1292).  mid real channels are stored 16-aligned, pads zero."""
    c0, c2, c4 = seq[0], seq[2], seq[4]
    mid = c0.weight.shape[0]
    midp = (mid + 15) // 16 * 16
    w0, b0 = packing.pack_conv(c0.weight, c0.bias, cin_pos, cin_width, midp)
    w2 = torch.zeros(midp, midp, device=device)
    w2[:mid, :mid] = c2.weight.detach().float().reshape(mid, mid)
    b2 = torch.zeros(midp, device=device)
    b2[:mid] = c2.bias.detach().float()
    w4, b4 = packing.pack_conv(c4.weight, c4.bias, torch.arange(mid, device=device), midp, 64)
    return dict(w0=w0, b0=b0, w2=w2.contiguous(), b2=b2, w4=w4, b4=b4, midp=midp)


class Executor:
    def __init__(self, module):
        self._module = weakref.ref(module)
        self._packed = None
        self._packed_key = None
        self._ws = {}
        self.last_workspace = None  # the workspace dict of the most recent forward (held by graph owners)
        self.use_tc = True          # bf16 mode: tcgen05 kernels (False = CUDA-core kernels on bf16 storage)

    def __deepcopy__(self, memo):
        return Executor.__new__(Executor)._reset()

    def _reset(self):
        self._module = lambda: None
        self._packed = self._packed_key = None
        self._ws = {}
        self.last_workspace = None
        self.use_tc = True
        return self

    def bound_to(self, module):
        return self._module() is module

    # ------------------------------------------------------------------ packed weights
    def _weights(self, device):
        m = self._module()
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in m.parameters())
        if self._packed is not None and key == self._packed_key:
            return self._packed
        with torch.no_grad():
            P = {"blocks": []}
            for blk in self._rdstbs(m):
                B = {"dstl": []}
                c = packing.EMBED
                for dstl in blk.body:
                    if getattr(dstl, "dim_modify_mode", "tail") == "head":
                        # LN + Linear(C -> 30) first, then the Swin blocks at width 30 (6 heads x 5): generic CUDA-core
                        # kernels (the tcgen05 kernels are built for 60 / 90 / 120)
                        B["dstl"].append(dict(
                            c=c, mode="head", head=packing.pack_dstl_head(dstl, c),
                            stl=[packing.pack_stl(b, packing.GROWTH) for b in dstl.body.blocks],
                            shifts=[b.shift_size for b in dstl.body.blocks], scale=float(m.dense_scale)))
                    else:
                        B["dstl"].append(dict(
                            c=c, mode="tail",
                            stl=[packing.pack_stl_tc(packing.pack_stl(b, c)) for b in dstl.body.blocks],
                            shifts=[b.shift_size for b in dstl.body.blocks],
                            tail=packing.pack_dstl_tail(dstl, c, m.dense_scale)))
                    c += packing.GROWTH
                pos = packing.channel_positions(c, device)
                if getattr(blk, "resi_connection", "1conv") == "3conv":
                    B["c3"] = _pack_3conv(blk.conv, pos, packing.DENSE_LD, device)     # 150 -> 37 (stored 48) -> 37 -> 60
                else:
                    B["lff_w"], B["lff_b"] = packing.pack_conv(blk.conv.weight, blk.conv.bias, pos, packing.DENSE_LD, 64)
                    B["lff_img"] = packing.conv_tc_image(B["lff_w"])
                P["blocks"].append(B)
            f = lambda t: t.detach().float().contiguous()
            id60 = torch.arange(60, device=device)
            P["head_w"] = f(m.head.weight).reshape(60, 9).contiguous()
            P["head_b"] = f(m.head.bias)
            P["pe_g"], P["pe_b"] = f(m.patch_embed.norm.weight), f(m.patch_embed.norm.bias)
            P["norm_g"], P["norm_b"] = f(m.norm.weight), f(m.norm.bias)
            if isinstance(m.conv_after_body, torch.nn.Sequential):
                P["cab3"] = _pack_3conv(m.conv_after_body, id60, 64, device)                 # 60 -> 15 (stored 16) -> 15 -> 60
                P["cab_w"] = P["cab_b"] = None
            else:
                P["cab_w"], P["cab_b"] = packing.pack_conv(m.conv_after_body.weight, m.conv_after_body.bias, id60, 64, 64)
            P["up"] = [packing.pack_upconv(l.weight, l.bias) for l in m.tail[0] if isinstance(l, torch.nn.Conv2d)]
            P["cab_img"] = packing.conv_tc_image(P["cab_w"]) if P["cab_w"] is not None else None
            P["up_img"] = [packing.conv_tc_image(w) for w, _ in P["up"]]
            last = m.tail[1]
            lw = last.weight.new_zeros(9, 64, dtype=torch.float32)
            lw[:, :60] = last.weight.detach().float()[0].permute(1, 2, 0).reshape(9, 60)
            P["last_w"] = lw.contiguous()
            P["last_img"] = packing.last_conv_tc_image(P["last_w"])
            # scalars are read once per re-pack (a host sync only when weights changed)
            P["last_b"] = float(last.bias.detach().float()[0])
            P["in_scale"] = float(m.sub_mean.weight.detach().reshape(-1)[0])
            P["in_bias"] = float(m.sub_mean.bias.detach().reshape(-1)[0])
            P["out_scale"] = float(m.add_mean.weight.detach().reshape(-1)[0])
            P["out_bias"] = float(m.add_mean.bias.detach().reshape(-1)[0])
        self._packed, self._packed_key = P, key
        return P

    # ------------------------------------------------------------------ workspaces
    def _workspace(self, B, H, W, dtype, device, scale):
        key = (B, H, W, dtype, str(device), scale)
        ws = self._ws.get(key)
        if ws is None:
            T = B * H * W
            e = lambda *s: torch.empty(*s, dtype=dtype, device=device)
            ws = dict(D=[e(T, 160), e(T, 160)], X1=e(T, 128), Y0=e(T, 128), Y1=e(T, 128),
                      QKV=e(T, 360), O=e(T, 120), HID=e(T, 240),
                      F0=e(T, 64), F1=e(T, 64),
                      FN=torch.zeros(T, 64, dtype=dtype, device=device))   # LN writes 60 ch; pads stay 0
            ups, t = [], T
            s = scale
            while s > 1:
                t *= 4
                ups.append(e(t, 64))
                s //= 2
            ws["UP"] = ups
            # Bounded cache: drop only the oldest shape.  A CUDA graph captured over a workspace keeps its own reference
            # to the dict (infer.GraphedRDST stores `last_workspace`), so eviction here can never free memory that a live
            # graph still addresses; every auxiliary buffer of the subclasses lives INSIDE the dict for the same reason.
            while len(self._ws) > 8:
                self._ws.pop(next(iter(self._ws)))
            self._ws[key] = ws
        self.last_workspace = ws
        return ws

    # ------------------------------------------------------------------ forward
    def forward(self, x, out=None):
        """out: optional fp32 (B,1,sH,sW) tensor the reconstruction kernel writes into -- a device tensor, or PINNED HOST
        memory (mapped into the device address space): the HR image then crosses PCIe as it is produced instead of in a
        separate device->host copy.  Inference only."""
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
            if out is not None:
                raise ValueError("rdst_b200: out= is an inference option (call under torch.no_grad())")
            from . import autograd
            return autograd.forward_with_grad(self, x)
        if out is not None:
            s = m.sr_scale
            ok = (out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (B, 1, H * s, W * s) and
                  (out.is_cuda and out.device == x.device or (not out.is_cuda and out.is_pinned())))
            if not ok:
                raise ValueError("rdst_b200: out= must be a contiguous fp32 (B,1,sH,sW) tensor on the input's device or in "
                                 "pinned host memory")
        with torch.no_grad(), torch.cuda.device(x.device):
            return self._forward_impl(x, out)

    def _forward_impl(self, x, out=None):
        m = self._module()
        dev = x.device
        adt = torch.float32 if m.precision == "fp32" else torch.bfloat16
        dt = _lib.dtype_code(adt)
        B, _, H, W = x.shape
        T = B * H * W
        P = self._weights(dev)
        ws = self._workspace(B, H, W, adt, dev, m.sr_scale)
        st = _lib.stream_ptr()
        xin = x.detach().to(torch.float32).contiguous()
        D = ws["D"]
        cur = 0
        call("rdst_head_fwd", ptr(xin), P["in_scale"], P["in_bias"], ptr(P["head_w"]), ptr(P["head_b"]),
             ptr(P["pe_g"]), ptr(P["pe_b"]), ptr(ws["F0"]), 64, ptr(D[cur]), 160, B, H, W, dt, st)
        for bi, blk in enumerate(P["blocks"]):
            self._block_start(bi, D[cur], T, ws)
            for j, ds in enumerate(blk["dstl"]):
                src, lds = D[cur], 160
                off = 64 + 32 * j
                if ds["mode"] == "head":
                    self._dstl_head_mode(ds, D[cur], off, B, H, W, ws, dt, st)
                    continue
                t = ds["tail"]
                fuse_tail = dt == _lib.BF16 and self.use_tc
                for k, (w, shift) in enumerate(zip(ds["stl"], ds["shifts"])):
                    cp = w["cp"]
                    dst = (ws["Y0"] if k == 0 else ws["Y1"]).view(-1)[:T * cp].view(T, cp)
                    last = k == len(ds["stl"]) - 1
                    self._stl(src, lds, dst, w, shift, B, H, W, ws, dt, st,
                              tail=(t, D[cur][:, off:]) if (fuse_tail and last) else None)
                    src, lds = dst, w["cp"]
                if not fuse_tail:
                    call("rdst_linear_fwd", ptr(src), lds, ptr(t["w"]), ptr(t["b"]), None, 0,
                         ptr(D[cur][:, off:]), 160, T, ds["stl"][0]["cp"], 32, ds["c"], 0, t["scale"], dt, st)
            if "c3" in blk:
                self._conv3(blk["c3"], D[cur], 160, 160, D[cur], 160, D[1 - cur], 160, B, H, W, float(m.rdb_residual_scale), ws, dt, st)
            else:
                self._conv(D[cur], 160, blk["lff_w"], blk["lff_img"], blk["lff_b"], D[cur], 160, D[1 - cur], 160,
                           B, H, W, 160, 64, float(m.rdb_residual_scale), 0, dt, st)
            cur = 1 - cur
            self._block_done(bi, D[cur], dict(P=P, ws=ws, B=B, H=H, W=W, T=T, dt=dt, st=st))
        feat = self._deep_features(D[cur], P, ws, B, H, W, T, dt, st)
        h, w_ = H, W
        for (uw, ub), uimg, buf in zip(P["up"], P["up_img"], ws["UP"]):
            self._conv(feat, 64, uw, uimg, ub, None, 0, buf, 64, B, h, w_, 64, 256, 1.0, 2, dt, st)
            feat, h, w_ = buf, 2 * h, 2 * w_
        given = out is not None
        if not given:
            out = torch.empty(B, 1, h, w_, dtype=torch.float32, device=dev)
        if dt == _lib.BF16 and self.use_tc:
            call("rdst_last_conv_fwd_bf16_tc", ptr(feat), 64, ptr(P["last_img"]), P["last_b"], P["out_scale"],
                 P["out_bias"], ptr(out), B, h, w_, st)
        else:
            call("rdst_last_conv_fwd", ptr(feat), 64, ptr(P["last_w"]), P["last_b"], P["out_scale"], P["out_bias"],
                 ptr(out), B, h, w_, 64, dt, st)
        return out if (given or x.dtype == torch.float32) else out.to(x.dtype)

    def _conv3(self, c3, x, ldx, cin, r, ldr, y, ldy, B, H, W, scale, ws, dt, st):
        """y = scale * conv3x3(lrelu(conv1x1(lrelu(conv3x3(x))))) + r  -- the '3conv' fusion on the generic CUDA-core kernels."""
        T, mp = B * H * W, c3["midp"]
        key = ("c3", mp)
        if key not in ws:
            ws[key] = [torch.empty(T, mp, dtype=x.dtype, device=x.device) for _ in range(2)]
        ta, tb = ws[key]
        call("rdst_conv3x3_act_fwd", ptr(x), ldx, ptr(c3["w0"]), ptr(c3["b0"]), ptr(ta), mp, B, H, W, cin, mp, 2, dt, st)
        call("rdst_linear_fwd", ptr(ta), mp, ptr(c3["w2"]), ptr(c3["b2"]), None, 0, ptr(tb), mp, T, mp, mp, 0, 2, 1.0, dt, st)
        call("rdst_conv3x3_fwd", ptr(tb), mp, ptr(c3["w4"]), ptr(c3["b4"]), ptr(r), ldr, ptr(y), ldy,
             B, H, W, mp, 64, scale, 0, dt, st)

    def _rdstbs(self, m):
        """The RDSTB modules of the network in execution order."""
        return list(m.body)

    def _block_start(self, index, trunk, T, ws):
        """Hook before RDSTB `index` reads the trunk."""

    def _block_done(self, index, trunk, c):
        """Hook after RDSTB `index` (trunk = dense buffer whose first 64 columns hold the block output; c = launch context)."""

    def _deep_features(self, trunk, P, ws, B, H, W, T, dt, st):
        """norm * global_res_scale -> conv_after_body -> + head output  (rdst_variations.py:1337-1350); returns the map
        that feeds the up-sampler."""
        m = self._module()
        call("rdst_layernorm_fwd", ptr(trunk), 160, ptr(P["norm_g"]), ptr(P["norm_b"]), ptr(ws["FN"]), 64,
             T, 60, float(m.global_res_scale), dt, st)
        if m.feature_last_operation and P.get("cab3") is not None:
            self._conv3(P["cab3"], ws["FN"], 64, 64, ws["F0"], 64, ws["F1"], 64, B, H, W, 1.0, ws, dt, st)
        elif m.feature_last_operation:
            self._conv(ws["FN"], 64, P["cab_w"], P["cab_img"], P["cab_b"], ws["F0"], 64, ws["F1"], 64,
                       B, H, W, 64, 64, 1.0, 0, dt, st)
        else:
            torch.add(ws["FN"], ws["F0"], out=ws["F1"])
        return ws["F1"]

    def _conv(self, x, ldx, w, wimg, b, r, ldr, y, ldy, B, H, W, cin, n, scale, shuffle, dt, st):
        if dt == _lib.BF16 and self.use_tc:
            call("rdst_conv3x3_fwd_bf16_tc", ptr(x), ldx, ptr(wimg), ptr(b), ptr(r), ldr, ptr(y), ldy,
                 B, H, W, cin, n, scale, shuffle, st)
        else:
            call("rdst_conv3x3_fwd", ptr(x), ldx, ptr(w), ptr(b), ptr(r), ldr, ptr(y), ldy,
                 B, H, W, cin, n, scale, shuffle, dt, st)

    def _dstl_head_mode(self, ds, dense, off, B, H, W, ws, dt, st):
        """DenseSTLayer in 'head' mode (rdst_variations.py:288-295, :335-340): growth = body(Linear(LN(x))) * dense_scale.
        The Swin blocks at width 30 run on the generic CUDA-core kernels with fp32 intermediates in BOTH precision modes
        (in bf16 mode only the dense buffer is bf16: storing qkv / hidden activations of this narrow path in bf16 as well
        costs ~0.03 dB, measured)."""
        T = B * H * W
        c, hd = ds["c"], ds["head"]
        cpi = packing.padded_width(c)
        key = ("head_mode", T)
        if key not in ws:
            f = lambda n: torch.empty(T, n, dtype=torch.float32, device=dense.device)
            ws[key] = dict(xin=f(128), h0=f(32), mid=f(32), out=f(32), QKV=f(96), O=f(32), X1=f(32), HID=f(64))
        w32 = ws[key]
        if dense.dtype == torch.float32:
            xin, ldx = dense, 160
        else:
            xin, ldx = w32["xin"].view(-1)[:T * cpi].view(T, cpi), cpi
            xin.copy_(dense[:, :cpi])
        call("rdst_linear_fwd", ptr(xin), ldx, ptr(hd["w"]), ptr(hd["b"]), None, 0, ptr(w32["h0"]), 32,
             T, cpi, 32, c, 0, 1.0, _lib.F32, st)
        srcs = [w32["h0"], w32["mid"]]
        n = len(ds["stl"])
        for k, (w, shift) in enumerate(zip(ds["stl"], ds["shifts"])):
            dst = w32["out"] if k == n - 1 else srcs[(k + 1) % 2]
            self._stl(srcs[k % 2], 32, dst, w, shift, B, H, W, w32, _lib.F32, st, generic=True)
        if ds["scale"] == 1.0:
            dense[:, off:off + 32].copy_(w32["out"])
        else:
            dense[:, off:off + 32].copy_(w32["out"] * ds["scale"])

    def _stl(self, src, lds, dst, w, shift, B, H, W, ws, dt, st, tail=None, generic=False, ldd=None):
        """One Swin block: x1 = x + proj(attn(LN1 x)); y = x1 + fc2(gelu(fc1(LN2 x1)))."""
        T = B * H * W
        c, cp, hp = w["c"], w["cp"], w["hp"]
        ldd = cp if ldd is None else ldd
        v = lambda buf, ld: buf.view(-1)[:T * ld].view(T, ld)       # compact [T][ld] view of a max-sized buffer
        qkv, o, x1, hid = v(ws["QKV"], 3 * c), v(ws["O"], c), v(ws["X1"], cp), v(ws["HID"], hp)
        if dt == _lib.BF16 and self.use_tc and not generic:
            call("rdst_stl_attn_fwd_bf16", ptr(src), lds, ptr(x1), cp, ptr(w["wqkv_img"]), ptr(w["wproj_img"]),
                 ptr(w["bqkv_tc"]), ptr(w["bproj"]), ptr(w["table_tc"]), B, H, W, c, shift, st)
            if tail is None:
                call("rdst_stl_mlp_fwd_bf16", ptr(x1), cp, ptr(dst), cp, ptr(w["w1img"]), ptr(w["w2img"]),
                     ptr(w["b1"]), ptr(w["b2"]), T, c, 0, st)
            else:   # second block of a DenseSTLayer: its output only feeds the tail LN+Linear -> dense slice
                t, dslice = tail
                call("rdst_stl_mlp_tail_fwd_bf16", ptr(x1), cp, ptr(w["w1img"]), ptr(w["w2img"]), ptr(w["b1"]),
                     ptr(w["b2"]), ptr(t["wimg"]), ptr(t["b"]), ptr(dslice), 160, t["scale"], T, c, 0, st)
            return
        call("rdst_linear_fwd", ptr(src), lds, ptr(w["wqkv"]), ptr(w["bqkv"]), None, 0, ptr(qkv), 3 * c,
             T, cp, 3 * c, c, 0, 1.0, dt, st)
        call("rdst_window_attention_fwd", ptr(qkv), 3 * c, ptr(w["table"]), ptr(o), c,
             B, H, W, c, packing.HEADS, shift, dt, st)
        call("rdst_linear_fwd", ptr(o), c, ptr(w["wproj"]), ptr(w["bproj"]), ptr(src), lds, ptr(x1), cp,
             T, c, cp, 0, 0, 1.0, dt, st)
        call("rdst_linear_fwd", ptr(x1), cp, ptr(w["w1"]), ptr(w["b1"]), None, 0, ptr(hid), hp,
             T, cp, hp, c, 1, 1.0, dt, st)
        call("rdst_linear_fwd", ptr(hid), hp, ptr(w["w2"]), ptr(w["b2"]), ptr(x1), cp, ptr(dst), ldd,
             T, hp, cp, 0, 0, 1.0, dt, st)


class ExecutorN(Executor):
    """RDSTSR_N (rdst_variations.py:824-1112) with the global bottleneck in 'mlp' mode: the outputs of all RDSTBs are
    concatenated channel-wise and reduced by two Linears ('mlp', :1071-1079) or a 1x1 + a 3x3 conv ('conv', :1080-1082); `norm` / `conv_after_body` are not on its path."""

    def _weights(self, device):
        fresh = self._packed is None
        P = super()._weights(device)
        if fresh or "bn_w1" not in P:
            m = self._module()
            n = len(m.body)
            with torch.no_grad():
                f = lambda t: t.detach().float()
                pos = torch.cat([torch.arange(60, device=device) + 64 * i for i in range(n)])     # real channel -> cat column
                w1 = torch.zeros(64, 64 * n, device=device)
                w1[:60, pos] = f(m.bottleneck[0].weight).reshape(60, 60 * n)      # Linear weight or 1x1 conv filter
                b1 = torch.zeros(64, device=device)
                b1[:60] = f(m.bottleneck[0].bias)
                P["bn_w1"], P["bn_b1"] = w1.contiguous(), b1
                if m.global_bottleneck_mode == "conv":
                    id60 = torch.arange(60, device=device)
                    P["bn_w2"], P["bn_b2"] = packing.pack_conv(m.bottleneck[1].weight, m.bottleneck[1].bias, id60, 64, 64)
                    P["bn_img2"] = packing.conv_tc_image(P["bn_w2"])
                else:
                    w2 = torch.zeros(64, 64, device=device)
                    w2[:60, :60] = f(m.bottleneck[1].weight)
                    b2 = torch.zeros(64, device=device)
                    b2[:60] = f(m.bottleneck[1].bias)
                    P["bn_w2"], P["bn_b2"] = w2.contiguous(), b2
        return P

    def _block_done(self, index, trunk, c):
        T, ws = c["T"], c["ws"]
        n = len(self._module().body)
        cat = ws.get("cat")                # per-workspace (shape, dtype, device) buffer: lives as long as the workspace
        if cat is None:
            cat = ws["cat"] = torch.empty(T, 64 * n, dtype=trunk.dtype, device=trunk.device)
        cat[:, 64 * index:64 * index + 64].copy_(trunk[:, :64])       # the reference's torch.cat of the RDSTB outputs

    def _deep_features(self, trunk, P, ws, B, H, W, T, dt, st):
        m = self._module()
        n = len(m.body)
        tmp = ws["FN"]
        call("rdst_linear_fwd", ptr(ws["cat"]), 64 * n, ptr(P["bn_w1"]), ptr(P["bn_b1"]), None, 0, ptr(tmp), 64,
             T, 64 * n, 64, 0, 0, 1.0, dt, st)
        if m.global_bottleneck_mode == "conv":
            self._conv(tmp, 64, P["bn_w2"], P["bn_img2"], P["bn_b2"], ws["F0"], 64, ws["F1"], 64,
                       B, H, W, 64, 64, float(m.global_res_scale), 0, dt, st)
        else:
            call("rdst_linear_fwd", ptr(tmp), 64, ptr(P["bn_w2"]), ptr(P["bn_b2"]), ptr(ws["F0"]), 64, ptr(ws["F1"]), 64,
                 T, 64, 64, 0, 0, float(m.global_res_scale), dt, st)
        return ws["F1"]


class ExecutorE(Executor):
    """ESTSR (rdst_variations.py:574-822): RDSTBs grouped into RRDSTBs, each group closed by a 3x3 conv, * rrdb_residual_scale,
    + the group's input (:548-555); the deep features are norm * global_res_scale + head output (no conv_after_body)."""

    def _rdstbs(self, m):
        return [r for rr in m.body for r in rr.body]

    def _weights(self, device):
        fresh = self._packed is None
        P = super()._weights(device)
        if fresh or "rr_conv" not in P:
            m = self._module()
            with torch.no_grad():
                id60 = torch.arange(60, device=device)
                P["rr_conv"], P["rr_end"], P["rr_start"], i = [], {}, set(), 0
                for g, rr in enumerate(m.body):
                    w, b = packing.pack_conv(rr.conv.weight, rr.conv.bias, id60, 64, 64)
                    P["rr_conv"].append((w, packing.conv_tc_image(w), b, float(rr.residual_scale)))
                    P["rr_start"].add(i)
                    i += len(rr.body)
                    P["rr_end"][i - 1] = g
        return P

    def _block_start(self, index, trunk, T, ws):
        if index in self._packed["rr_start"]:            # keep the group's input: the RRDSTB shortcut
            if "rr_short" not in ws:                     # own buffers (the Swin-block workspaces are in use inside a group),
                for k in ("rr_short", "rr_tmp", "rr_in"):        # owned by the workspace dict so that they share its lifetime
                    ws[k] = torch.empty(T, 64, dtype=trunk.dtype, device=trunk.device)
            ws["rr_short"].copy_(trunk[:, :64])

    def _block_done(self, index, trunk, c):
        P, ws = c["P"], c["ws"]
        if index in P["rr_end"]:
            w, img, b, scale = P["rr_conv"][P["rr_end"][index]]
            ws["rr_in"].copy_(trunk[:, :64])             # compact [T][64] map, the layout conv_after_body runs on
            self._conv(ws["rr_in"], 64, w, img, b, ws["rr_short"], 64, ws["rr_tmp"], 64, c["B"], c["H"], c["W"], 64, 64, scale, 0,
                       c["dt"], c["st"])
            trunk[:, :64].copy_(ws["rr_tmp"])
