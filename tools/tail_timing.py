"""Stand-alone timings of the reconstruction tail kernels at the cfg2 shape (up-convs and the final conv)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
from rdst_b200 import _lib, packing
g = torch.Generator().manual_seed(0)
B = 176
def timeit(fn, n=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (H, W) in ((40, 32), (80, 64)):
    T = B * H * W
    x = torch.randn(T, 64, generator=g).to(torch.bfloat16).cuda()
    w = torch.randn(256, 9, 64, generator=g) * 0.05
    img = packing.conv_tc_image(w).cuda()
    bias = torch.zeros(256).cuda()
    y = torch.empty(T * 4, 64, dtype=torch.bfloat16, device="cuda")
    us = timeit(lambda: _lib.call("rdst_conv3x3_fwd_bf16_tc", _lib.ptr(x), 64, _lib.ptr(img), _lib.ptr(bias), None, 0, _lib.ptr(y), 64,
                                  B, H, W, 64, 256, 1.0, 2, _lib.stream_ptr()))
    mb = (T * 128 + T * 4 * 128) / 1e6
    print(f"up-conv 64->256 + shuffle at {H}x{W}: {us:7.1f} us  ({mb / us * 1e-3 * 1e3:.0f} GB/s algorithmic)")
H, W = 160, 128
T = B * H * W
x = torch.randn(T, 64, generator=g).to(torch.bfloat16).cuda()
lw = torch.randn(9, 64, generator=g) * 0.05
limg = packing.last_conv_tc_image(lw).cuda()
out = torch.empty(B, 1, H, W, device="cuda")
us = timeit(lambda: _lib.call("rdst_last_conv_fwd_bf16_tc", _lib.ptr(x), 64, _lib.ptr(limg), 0.1, 1.0, 0.0, _lib.ptr(out), B, H, W, _lib.stream_ptr()))
mb = (T * 128 + T * 4) / 1e6
print(f"last conv 64->1 at {H}x{W}: {us:7.1f} us  ({mb / us:.0f} GB/s algorithmic)")
