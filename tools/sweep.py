"""BASELINE cfg5: large-input sweep (RDST-E1 x2/x4, LR 128..512 px) + window-attention micro-benchmark
(C in {60,90,120}, shifted vs non-shifted).  Device-side timing with CUDA events; prints a table."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers
from rdst_b200 import _lib, packing


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("== whole network, bf16, batch 1 ==")
for scale in (4, 2):
    m = helpers.make_module(8, scale, "bf16").cuda().eval()
    for s in (128, 256, 384, 512):
        x = torch.rand(1, 1, s, s, device="cuda")
        with torch.no_grad():
            ms = timeit(lambda: m(x))
            y = m(x)
        assert torch.isfinite(y).all() and y.shape == (1, 1, s * scale, s * scale)
        print(f"E1 x{scale} LR {s}x{s}: {ms:8.2f} ms  {s * s * scale * scale / ms / 1e3:9.1f} HR Mpix/s")

print("== fused window attention kernel: shifted vs non-shifted ==")
g = torch.Generator().manual_seed(0)
for c in (60, 90, 120):
    cp = packing.padded_width(c)
    pos = packing.channel_positions(c)
    wqkv = torch.zeros(3 * c, cp); wqkv[:, pos] = torch.randn(3 * c, c, generator=g) * 0.1
    wproj = torch.zeros(cp, c); wproj[pos] = torch.randn(c, c, generator=g) * 0.1
    pk = {k: v.cuda() for k, v in packing.pack_attn_tc(wqkv, torch.zeros(3 * c), wproj, torch.zeros(cp),
                                                       torch.randn(225, 6, generator=g), c).items()}
    bp = torch.zeros(cp).cuda()
    for nw, (B, H, W) in ((256, (1, 128, 128)), (1024, (1, 256, 256)), (4096, (1, 512, 512))):
        T = B * H * W
        x = torch.zeros(T, cp); x[:, pos] = torch.randn(T, c, generator=g)
        x = x.to(torch.bfloat16).cuda()
        y = torch.empty_like(x)
        res = []
        for shift in (0, 4):
            ms = timeit(lambda: _lib.call("rdst_stl_attn_fwd_bf16", _lib.ptr(x), cp, _lib.ptr(y), cp, _lib.ptr(pk["wqkv_img"]),
                                          _lib.ptr(pk["wproj_img"]), _lib.ptr(pk["bqkv_tc"]), _lib.ptr(bp),
                                          _lib.ptr(pk["table_tc"]), B, H, W, c, shift, _lib.stream_ptr()), n=10)
            flop = (512 * c * c + 16384 * c) * nw
            res.append((ms, flop / ms / 1e9))
        print(f"C={c:3d} nW={nw:5d}: unshifted {res[0][0] * 1e3:7.1f} us ({res[0][1]:6.1f} TFLOP/s)   "
              f"shifted {res[1][0] * 1e3:7.1f} us ({res[1][1]:6.1f} TFLOP/s)")
