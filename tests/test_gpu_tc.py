"""GPU parity of the tcgen05 kernels (UMMA self-test, fused MLP, fused window attention) against contract
restatements on seeded inputs."""
import pytest
import torch

import abi_emulator as E

pytestmark = pytest.mark.gpu


def _L():
    from rdst_b200 import _lib
    return _lib


def _rand(shape, seed, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


@pytest.mark.parametrize("N,K,mn", [(16, 16, 0), (64, 64, 0), (240, 128, 0), (128, 240, 0), (32, 128, 1), (16, 128, 1),
                                    (128, 32, 0), (256, 256, 1)])
def test_umma_selftest(N, K, mn):
    L = _L()
    A = _rand((128, K), 1, 0.5).to(torch.bfloat16).cuda()
    B = _rand((N, K), 2, 0.5).to(torch.bfloat16).cuda()
    D = torch.zeros(128, N, device="cuda")
    L.call("rdst_umma_selftest", L.ptr(A), L.ptr(B), L.ptr(D), N, K, mn, 0, L.stream_ptr())
    ref = A.float() @ B.float().t()
    assert (D - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("C,ld,shift,hs0,ws0", [(64, 64, 0, 8, 16), (96, 160, 4, 32, 24), (128, 128, 4, 16, 24),
                                                 (96, 96, 0, 0, 0), (128, 160, 4, 32, 0)])
def test_tma_selftest(C, ld, shift, hs0, ws0):
    """4x4-token TMA boxes of a (wrapped) shifted window: SWIZZLE_128B shared-memory image and the store round trip."""
    L = _L()
    B, H, W, b = 3, 40, 32, 1
    x = _rand((B, H, W, ld), 21).to(torch.bfloat16).cuda()
    y = torch.full((B, H, W, 128), 7.0, device="cuda", dtype=torch.bfloat16)
    npan = (C + 63) // 64
    dump = torch.zeros(npan * 4096, device="cuda", dtype=torch.bfloat16)
    L.call("rdst_tma_selftest", L.ptr(x), ld, L.ptr(y), 128, B, H, W, C, shift, b, hs0, ws0, L.ptr(dump), L.stream_ptr())
    torch.cuda.synchronize()
    got = dump.cpu().view(npan, 64, 8, 8)
    xc = x.cpu()
    exp = torch.zeros(npan, 64, 8, 8, dtype=torch.bfloat16)
    yy = torch.full((B, H, W, 128), 7.0, dtype=torch.bfloat16)
    for iy in range(8):
        for ix in range(8):
            r = ((iy >> 2) * 2 + (ix >> 2)) * 16 + (iy & 3) * 4 + (ix & 3)      # box-major row order
            h, w = (hs0 + iy + shift) % H, (ws0 + ix + shift) % W
            v = torch.zeros(npan * 64, dtype=torch.bfloat16)
            v[:C] = xc[b, h, w, :C]
            for p in range(npan):
                for c in range(8):
                    exp[p, r, c ^ (r & 7)] = v[p * 64 + c * 8: p * 64 + c * 8 + 8]
            yy[b, h, w, :C] = xc[b, h, w, :C]
    assert (got == exp).all()
    assert (y.cpu() == yy).all()


def _padded_input(T, c, seed):
    from rdst_b200 import packing
    cp = packing.padded_width(c)
    x = torch.zeros(T, cp)
    x[:, packing.channel_positions(c)] = _rand((T, c), seed)
    return x.to(torch.bfloat16), cp


# the last three cases have more tiles than SMs (148): the persistent multi-tile path, incl. a ragged last tile
@pytest.mark.parametrize("c,T", [(60, 128), (60, 1000), (90, 640), (120, 128), (120, 19 * 128 + 64),
                                 (60, 128 * 333 + 17), (90, 128 * 401 + 64), (120, 128 * 590 + 100)])
@pytest.mark.parametrize("exact", [0, 1])
def test_fused_mlp(c, T, exact):
    from rdst_b200 import packing
    L = _L()
    x, cp = _padded_input(T, c, 3)
    hp = packing.hidden_width(2 * c)
    pos = packing.channel_positions(c)
    w1 = torch.zeros(hp, cp); w1[:2 * c, pos] = _rand((2 * c, c), 4, 0.08)
    b1 = torch.zeros(hp); b1[:2 * c] = _rand((2 * c,), 5, 0.1)
    w2 = torch.zeros(cp, hp); w2[pos, :2 * c] = _rand((c, 2 * c), 6, 0.08)
    b2 = torch.zeros(cp); b2[pos] = _rand((c,), 7, 0.1)
    w1b, w2b = w1.to(torch.bfloat16).float(), w2.to(torch.float16).float()
    # contract restatement (fp32 math on the same bf16-rounded weights / inputs)
    hid = torch.zeros(T, hp)
    E.rdst_linear_fwd(x, cp, w1b, b1, None, 0, hid, hp, T, cp, hp, c, 1, 1.0, 0, None)
    ref = torch.zeros(T, cp)
    E.rdst_linear_fwd(hid, hp, w2b, b2, x, cp, ref, cp, T, hp, cp, 0, 0, 1.0, 0, None)
    xd = x.cuda()
    yd = torch.full((T, cp), float("nan"), dtype=torch.bfloat16, device="cuda")
    dev = [packing.fc1_image(w1).cuda(), packing.fc2_image(w2).cuda(), b1.cuda(), b2.cuda()]
    L.call("rdst_stl_mlp_fwd_bf16", L.ptr(xd), cp, L.ptr(yd), cp, L.ptr(dev[0]), L.ptr(dev[1]), L.ptr(dev[2]),
           L.ptr(dev[3]), T, c, exact, L.stream_ptr())
    y = yd.cpu().float()
    assert torch.isfinite(y).all()
    err = (y - ref).abs()
    assert err.max().item() < 6e-2 and err.mean().item() < 6e-3, (err.max().item(), err.mean().item())
    padmask = torch.ones(cp, dtype=torch.bool); padmask[pos] = False
    assert (y[:, padmask] == 0).all()            # pads stay exactly zero


@pytest.mark.parametrize("c,B,H,W,shift", [(60, 2, 16, 24, 0), (60, 2, 16, 24, 4), (90, 1, 24, 24, 4), (120, 1, 8, 8, 4),
                                            (120, 3, 40, 32, 4), (120, 3, 40, 32, 0), (90, 5, 8, 16, 0),
                                            # more tiles than SMs (148), shifted and not, odd window counts: the
                                            # persistent multi-tile path (next-tile prefetch, C=120 landing zone)
                                            (120, 43, 40, 24, 4), (120, 43, 40, 24, 0), (90, 37, 24, 40, 4),
                                            (90, 37, 24, 40, 0), (60, 51, 24, 24, 4), (60, 51, 24, 24, 0),
                                            (120, 1, 264, 264, 4), (90, 1, 264, 264, 4)])
def test_fused_attention(c, B, H, W, shift):
    from rdst_b200 import packing
    L = _L()
    T = B * H * W
    x, cp = _padded_input(T, c, 11)
    pos = packing.channel_positions(c)
    hd = c // 6
    wqkv = torch.zeros(3 * c, cp); wqkv[:, pos] = _rand((3 * c, c), 12, 0.12)
    bqkv = _rand((3 * c,), 13, 0.2)
    wqkv[:c] *= hd ** -0.5; bqkv[:c] *= hd ** -0.5
    wproj = torch.zeros(cp, c); wproj[pos] = _rand((c, c), 14, 0.1)
    bproj = torch.zeros(cp); bproj[pos] = _rand((c,), 15, 0.1)
    table = _rand((225, 6), 16, 0.7)
    rb = lambda t: t.to(torch.bfloat16).float()
    # contract restatement on bf16-rounded weights
    qkv = torch.zeros(T, 3 * c)
    E.rdst_linear_fwd(x, cp, rb(wqkv), bqkv, None, 0, qkv, 3 * c, T, cp, 3 * c, c, 0, 1.0, 0, None)
    o = torch.zeros(T, c)
    E.rdst_window_attention_fwd(qkv, 3 * c, table, o, c, B, H, W, c, 6, shift, 0, None)
    ref = torch.zeros(T, cp)
    E.rdst_linear_fwd(o, c, rb(wproj), bproj, x, cp, ref, cp, T, c, cp, 0, 0, 1.0, 0, None)
    pk = packing.pack_attn_tc(wqkv, bqkv, wproj, bproj, table, c)
    xd = x.cuda()
    yd = torch.full((T, cp), float("nan"), dtype=torch.bfloat16, device="cuda")
    d = {k: v.cuda() for k, v in pk.items()}
    bp = bproj.cuda()
    L.call("rdst_stl_attn_fwd_bf16", L.ptr(xd), cp, L.ptr(yd), cp, L.ptr(d["wqkv_img"]), L.ptr(d["wproj_img"]),
           L.ptr(d["bqkv_tc"]), L.ptr(bp), L.ptr(d["table_tc"]), B, H, W, c, shift, L.stream_ptr())
    y = yd.cpu().float()
    assert torch.isfinite(y).all()
    err = (y - ref).abs()
    assert err.max().item() < 6e-2 and err.mean().item() < 6e-3, (err.max().item(), err.mean().item())
    padmask = torch.ones(cp, dtype=torch.bool); padmask[pos] = False
    assert (y[:, padmask] == 0).all()


@pytest.mark.parametrize("B,H,W,Cin,N,shuffle,res", [(2, 8, 16, 160, 64, 0, True), (3, 40, 32, 160, 64, 0, True),
                                                      (1, 24, 24, 64, 64, 0, True), (2, 8, 8, 64, 256, 2, False),
                                                      (1, 80, 64, 64, 256, 2, False), (1, 16, 40, 64, 64, 0, False),
                                                      (1, 64, 64, 160, 64, 0, True)])
def test_conv3x3_tc(B, H, W, Cin, N, shuffle, res):
    from rdst_b200 import packing
    L = _L()
    T = B * H * W
    x = _rand((T, Cin), 21).to(torch.bfloat16)
    w, b = _rand((N, 9, Cin), 22, 0.05), _rand((N,), 23, 0.1)
    To, ldy = (4 * T, N // 4) if shuffle else (T, N)
    r = _rand((T, ldy), 24).to(torch.bfloat16) if res else None
    ref = torch.zeros(To, ldy)
    E.rdst_conv3x3_fwd(x, Cin, w.to(torch.bfloat16).float(), b, r, ldy, ref, ldy, B, H, W, Cin, N, 0.75, shuffle, 0, None)
    xd, bd, wimg = x.cuda(), b.cuda(), packing.conv_tc_image(w).cuda()
    rd = r.cuda() if res else None
    yd = torch.full((To, ldy), float("nan"), dtype=torch.bfloat16, device="cuda")
    L.call("rdst_conv3x3_fwd_bf16_tc", L.ptr(xd), Cin, L.ptr(wimg), L.ptr(bd), L.ptr(rd), ldy, L.ptr(yd), ldy,
           B, H, W, Cin, N, 0.75, shuffle, L.stream_ptr())
    y = yd.cpu().float()
    assert torch.isfinite(y).all()
    err = (y - ref).abs()
    assert err.max().item() < 8e-2 and err.mean().item() < 8e-3, (err.max().item(), err.mean().item())


@pytest.mark.parametrize("c,T", [(60, 640), (90, 128 * 3 + 64), (120, 1280)])
def test_fused_mlp_with_dense_tail(c, T):
    from rdst_b200 import packing
    L = _L()
    x, cp = _padded_input(T, c, 31)
    hp = packing.hidden_width(2 * c)
    pos = packing.channel_positions(c)
    w1 = torch.zeros(hp, cp); w1[:2 * c, pos] = _rand((2 * c, c), 32, 0.08)
    b1 = torch.zeros(hp); b1[:2 * c] = _rand((2 * c,), 33, 0.1)
    w2 = torch.zeros(cp, hp); w2[pos, :2 * c] = _rand((c, 2 * c), 34, 0.08)
    b2 = torch.zeros(cp); b2[pos] = _rand((c,), 35, 0.1)
    wt = torch.zeros(32, cp); wt[:30, pos] = _rand((30, c), 36, 0.1)
    bt = torch.zeros(32); bt[:30] = _rand((30,), 37, 0.1)
    rb = lambda t: t.to(torch.bfloat16).float()
    hid = torch.zeros(T, hp)
    E.rdst_linear_fwd(x, cp, rb(w1), b1, None, 0, hid, hp, T, cp, hp, c, 1, 1.0, 0, None)
    y = torch.zeros(T, cp)
    E.rdst_linear_fwd(hid, hp, rb(w2), b2, x, cp, y, cp, T, hp, cp, 0, 0, 1.0, 0, None)
    ref = torch.zeros(T, 32)
    E.rdst_linear_fwd(y, cp, rb(wt), bt, None, 0, ref, 32, T, cp, 32, c, 0, 0.5, 0, None)
    dense = torch.full((T, 160), 7.0, dtype=torch.bfloat16, device="cuda")
    dev = [t.cuda() for t in (x, packing.fc1_image(w1), packing.fc2_image(w2), b1, b2, packing.kmajor_image(wt), bt)]
    L.call("rdst_stl_mlp_tail_fwd_bf16", L.ptr(dev[0]), cp, L.ptr(dev[1]), L.ptr(dev[2]), L.ptr(dev[3]), L.ptr(dev[4]),
           L.ptr(dev[5]), L.ptr(dev[6]), L.ptr(dense[:, 96:]), 160, 0.5, T, c, 0, L.stream_ptr())
    d = dense.cpu().float()
    err = (d[:, 96:128] - ref).abs()
    assert err.max().item() < 6e-2 and err.mean().item() < 6e-3, (err.max().item(), err.mean().item())
    assert (d[:, :96] == 7.0).all() and (d[:, 128:] == 7.0).all()      # only the 32-wide slice is written
    assert (d[:, 126:128] == 0).all()


@pytest.mark.parametrize("B,H,W", [(2, 32, 48), (1, 8, 8), (3, 160, 128), (2, 50, 70), (1, 16, 200), (5, 96, 24)])
def test_last_conv_tc(B, H, W):
    """tap-GEMM reconstruction conv: tiles of up to four 128-position blocks, ragged right / bottom tiles, tiny images"""
    from rdst_b200 import packing
    L = _L()
    T = B * H * W
    x = _rand((T, 64), 41).to(torch.bfloat16)
    lw = torch.zeros(9, 64); lw[:, :60] = _rand((9, 60), 42, 0.1)
    ref = torch.zeros(B, 1, H, W)
    E.rdst_last_conv_fwd(x, 64, lw.to(torch.bfloat16).float(), 0.3, 2.0, 0.1, ref, B, H, W, 64, 0, None)
    xd, img = x.cuda(), packing.last_conv_tc_image(lw).cuda()
    od = torch.full((B, 1, H, W), float("nan"), device="cuda")
    L.call("rdst_last_conv_fwd_bf16_tc", L.ptr(xd), 64, L.ptr(img), 0.3, 2.0, 0.1, L.ptr(od), B, H, W, L.stream_ptr())
    assert (od.cpu() - ref).abs().max().item() < 2e-3
