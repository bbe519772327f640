"""Launch time of the reconstruction conv (rdst_last_conv_fwd_bf16_tc) at the cfg2 output size (176 x 160 x 128 x 64ch)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdst_b200 import _lib as L, packing
B, H, W = 176, 160, 128
x = torch.randn(B * H * W, 64, device="cuda").to(torch.bfloat16)
lw = torch.zeros(9, 64); lw[:, :60] = torch.randn(9, 60) * 0.1
img = packing.last_conv_tc_image(lw).cuda()
out = torch.empty(B, 1, H, W, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(): L.call("rdst_last_conv_fwd_bf16_tc", L.ptr(x), 64, L.ptr(img), 0.3, 2.0, 0.1, L.ptr(out), B, H, W, L.stream_ptr())
for _ in range(3): run()
ts = []
for _ in range(10):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
ts.sort()
gb = B * H * W * 128 / 1e9
print(f"last conv {B}x{H}x{W}: {sum(ts)/len(ts):.1f} us (min {ts[0]:.1f}) -> {gb / (ts[0] * 1e-6) / 1e3:.2f} TB/s of input")
