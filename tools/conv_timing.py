"""Phase-level cycle timeline of CTA 0 of the tcgen05 3x3 convolution (uses rdst_debug_conv_timing)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
from rdst_b200 import _lib, packing
cin = int(sys.argv[1]) if len(sys.argv) > 1 else 160
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
B, H, W = 176, 40, 32
T = B * H * W
g = torch.Generator().manual_seed(0)
x = (torch.randn(T, cin, generator=g)).to(torch.bfloat16).cuda()
w = torch.randn(n, 9, cin, generator=g) * 0.05          # [N][tap][Cin]
img = packing.conv_tc_image(w).cuda()
bias = torch.zeros(n).cuda()
shuffle = 2 if n == 256 else 0
y = torch.empty(T * (4 if shuffle else 1), 64, dtype=torch.bfloat16, device="cuda")
dbg = torch.zeros(128, dtype=torch.int64, device="cuda")
def run():
    _lib.call("rdst_conv3x3_fwd_bf16_tc", _lib.ptr(x), cin, _lib.ptr(img), _lib.ptr(bias), None, 0, _lib.ptr(y), 64,
              B, H, W, cin, n, 1.0, shuffle, _lib.stream_ptr())
run(); torch.cuda.synchronize()
_lib.call("rdst_debug_conv_timing", _lib.ptr(dbg)); run(); torch.cuda.synchronize(); _lib.call("rdst_debug_conv_timing", None)
d = dbg.cpu().tolist()
names = ["tile start", "next staged", "MMAs done", "epilogue done"]
for tile in range(2, 6):
    base = d[tile * 4]
    prev = base
    for k, nm in enumerate(names):
        v = d[tile * 4 + k]
        print(f"  tile{tile} {nm:16s} +{v - prev:6d}  (t={v - base:6d})")
        prev = v
    print(f"  tile{tile} -> next tile start +{d[tile * 4 + 4] - prev:6d}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print(f"Cin={cin} N={n}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
