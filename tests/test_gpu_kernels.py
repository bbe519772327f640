"""GPU parity of every C-ABI kernel against its contract restatement (tests/abi_emulator.py) on seeded inputs,
through the same ctypes entry points the module uses."""
import pytest
import torch

import abi_emulator as E

pytestmark = pytest.mark.gpu


def _lib():
    from rdst_b200 import _lib
    return _lib


def _tol(dtype):
    return 2e-4 if dtype == torch.float32 else 6e-2


def _rand(shape, dtype, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("T,K,N,creal,act,res", [(257, 64, 180, 60, 0, False), (300, 128, 240, 120, 1, False),
                                                  (129, 192, 96, 0, 0, True), (64, 90, 96, 0, 0, True),
                                                  (1000, 96, 32, 90, 0, False)])
def test_linear(dtype, T, K, N, creal, act, res):
    L = _lib()
    ldx, ldy = K + 32, N + 16
    x = _rand((T, ldx), dtype, 1)
    if creal:
        x[:, creal:K] = 0
    w, b = _rand((N, K), torch.float32, 2, 0.1), _rand((N,), torch.float32, 3, 0.1)
    r = _rand((T, ldy), dtype, 4) if res else None
    y_ref = torch.zeros(T, ldy, dtype=dtype)
    E.rdst_linear_fwd(x, ldx, w, b, r, ldy, y_ref, ldy, T, K, N, creal, act, 0.5, 0, None)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    rd = r.cuda() if res else None
    yd = torch.zeros(T, ldy, dtype=dtype, device="cuda")
    L.call("rdst_linear_fwd", L.ptr(xd), ldx, L.ptr(wd), L.ptr(bd), L.ptr(rd), ldy, L.ptr(yd), ldy, T, K, N, creal, act,
           0.5, L.dtype_code(dtype), L.stream_ptr())
    assert (yd.cpu().float() - y_ref.float()).abs().max().item() < _tol(dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,C,shift", [(2, 16, 24, 60, 0), (2, 16, 24, 60, 4), (1, 8, 8, 90, 4), (1, 40, 32, 120, 4),
                                            (3, 24, 24, 120, 0)])
def test_window_attention(dtype, B, H, W, C, shift):
    L = _lib()
    T = B * H * W
    qkv = _rand((T, 3 * C), dtype, 5)
    table = _rand((225, 6), torch.float32, 6, 0.5)
    o_ref = torch.zeros(T, C, dtype=dtype)
    E.rdst_window_attention_fwd(qkv, 3 * C, table, o_ref, C, B, H, W, C, 6, shift, 0, None)
    qd, td = qkv.cuda(), table.cuda()
    od = torch.zeros(T, C, dtype=dtype, device="cuda")
    L.call("rdst_window_attention_fwd", L.ptr(qd), 3 * C, L.ptr(td), L.ptr(od), C, B, H, W, C, 6, shift,
           L.dtype_code(dtype), L.stream_ptr())
    assert (od.cpu().float() - o_ref.float()).abs().max().item() < _tol(dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,Cin,N,shuffle,res", [(2, 8, 16, 160, 64, 0, True), (1, 16, 8, 64, 64, 0, True),
                                                      (2, 8, 8, 64, 256, 2, False), (1, 24, 40, 64, 256, 2, False)])
def test_conv3x3(dtype, B, H, W, Cin, N, shuffle, res):
    L = _lib()
    T = B * H * W
    x = _rand((T, Cin), dtype, 7)
    w, b = _rand((N, 9, Cin), torch.float32, 8, 0.05), _rand((N,), torch.float32, 9, 0.1)
    To, ldy = (4 * T, N // 4) if shuffle else (T, N)
    r = _rand((T, ldy), dtype, 10) if res else None
    y_ref = torch.zeros(To, ldy, dtype=dtype)
    E.rdst_conv3x3_fwd(x, Cin, w, b, r, ldy, y_ref, ldy, B, H, W, Cin, N, 0.75, shuffle, 0, None)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    rd = r.cuda() if res else None
    yd = torch.zeros(To, ldy, dtype=dtype, device="cuda")
    L.call("rdst_conv3x3_fwd", L.ptr(xd), Cin, L.ptr(wd), L.ptr(bd), L.ptr(rd), ldy, L.ptr(yd), ldy, B, H, W, Cin, N,
           0.75, shuffle, L.dtype_code(dtype), L.stream_ptr())
    assert (yd.cpu().float() - y_ref.float()).abs().max().item() < _tol(dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_head_layernorm_lastconv(dtype):
    L = _lib()
    B, H, W = 2, 16, 24
    T = B * H * W
    img = torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(11))
    w, b = _rand((60, 9), torch.float32, 12, 0.3), _rand((60,), torch.float32, 13, 0.1)
    g, be = 1 + _rand((60,), torch.float32, 14, 0.1), _rand((60,), torch.float32, 15, 0.1)
    f_ref, d_ref = torch.zeros(T, 64, dtype=dtype), torch.zeros(T, 160, dtype=dtype)
    E.rdst_head_fwd(img, 2.0, -0.25, w, b, g, be, f_ref, 64, d_ref, 160, B, H, W, 0, None)
    fd, dd = torch.zeros(T, 64, dtype=dtype, device="cuda"), torch.zeros(T, 160, dtype=dtype, device="cuda")
    dev = [t.cuda() for t in (img, w, b, g, be)]
    L.call("rdst_head_fwd", L.ptr(dev[0]), 2.0, -0.25, L.ptr(dev[1]), L.ptr(dev[2]), L.ptr(dev[3]), L.ptr(dev[4]),
           L.ptr(fd), 64, L.ptr(dd), 160, B, H, W, L.dtype_code(dtype), L.stream_ptr())
    assert (fd.cpu().float() - f_ref.float()).abs().max().item() < _tol(dtype)
    assert (dd.cpu().float() - d_ref.float()).abs().max().item() < _tol(dtype)
    # layernorm
    x = _rand((T, 160), dtype, 16)
    y_ref = torch.zeros(T, 64, dtype=dtype)
    E.rdst_layernorm_fwd(x, 160, g, be, y_ref, 64, T, 60, 0.5, 0, None)
    xd, yd = x.cuda(), torch.zeros(T, 64, dtype=dtype, device="cuda")
    L.call("rdst_layernorm_fwd", L.ptr(xd), 160, L.ptr(dev[3]), L.ptr(dev[4]), L.ptr(yd), 64, T, 60, 0.5,
           L.dtype_code(dtype), L.stream_ptr())
    assert (yd.cpu().float() - y_ref.float()).abs().max().item() < _tol(dtype)
    # last conv
    x = _rand((T, 64), dtype, 17)
    lw = _rand((9, 64), torch.float32, 18, 0.1)
    o_ref = torch.zeros(B, 1, H, W)
    E.rdst_last_conv_fwd(x, 64, lw, 0.3, 2.0, 0.1, o_ref, B, H, W, 64, 0, None)
    xd, lwd, od = x.cuda(), lw.cuda(), torch.zeros(B, 1, H, W, device="cuda")
    L.call("rdst_last_conv_fwd", L.ptr(xd), 64, L.ptr(lwd), 0.3, 2.0, 0.1, L.ptr(od), B, H, W, 64,
           L.dtype_code(dtype), L.stream_ptr())
    assert (od.cpu() - o_ref).abs().max().item() < 1e-3


def test_abi_errors_are_reported():
    L = _lib()
    x = torch.zeros(8, 8, device="cuda")
    with pytest.raises(RuntimeError, match="multiples of the window size"):
        L.call("rdst_window_attention_fwd", L.ptr(x), 180, L.ptr(x), L.ptr(x), 60, 1, 12, 16, 60, 6, 0, 0, L.stream_ptr())
    with pytest.raises(RuntimeError, match="null pointer"):
        L.call("rdst_linear_fwd", None, 8, L.ptr(x), None, None, 0, L.ptr(x), 8, 8, 8, 8, 0, 0, 1.0, 0, L.stream_ptr())
