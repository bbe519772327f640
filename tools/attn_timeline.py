"""Whole-launch tile timeline of the middle CTA of the fused attention kernel (debug mode 1)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
from rdst_b200 import _lib, packing
c = int(sys.argv[1]) if len(sys.argv) > 1 else 120
shift = int(sys.argv[2]) if len(sys.argv) > 2 else 0
B, H, W = 176, 40, 32
T = B * H * W
cp = packing.padded_width(c)
pos = packing.channel_positions(c)
g = torch.Generator().manual_seed(0)
x = torch.zeros(T, cp); x[:, pos] = torch.randn(T, c, generator=g)
x = x.to(torch.bfloat16).cuda()
wqkv = torch.zeros(3 * c, cp); wqkv[:, pos] = torch.randn(3 * c, c, generator=g) * 0.1
wproj = torch.zeros(cp, c); wproj[pos] = torch.randn(c, c, generator=g) * 0.1
pk = {k: v.cuda() for k, v in packing.pack_attn_tc(wqkv, torch.zeros(3 * c), wproj, torch.zeros(cp), torch.randn(225, 6, generator=g), c).items()}
bp = torch.zeros(cp).cuda()
y = torch.empty_like(x)
dbg = torch.zeros(128, dtype=torch.int64, device="cuda")
def run():
    _lib.call("rdst_stl_attn_fwd_bf16", _lib.ptr(x), cp, _lib.ptr(y), cp, _lib.ptr(pk["wqkv_img"]), _lib.ptr(pk["wproj_img"]),
              _lib.ptr(pk["bqkv_tc"]), _lib.ptr(bp), _lib.ptr(pk["table_tc"]), B, H, W, c, shift, _lib.stream_ptr())
for _ in range(3): run()
torch.cuda.synchronize()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
if os.environ.get("RDST_COLD"):
    flush.fill_(1)          # evict x / y / weights from L2: the launch below starts cold
    torch.cuda.synchronize()
import ctypes; _lib.call("rdst_debug_attn_timing", ctypes.c_void_p(dbg.data_ptr() + 1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
_lib.call("rdst_debug_attn_timing", None)
d = dbg.cpu().tolist()[:64]
d = [v for v in d if v]
print(f"C={c}: launch {e0.elapsed_time(e1) * 1e3:.1f} us (DBG build); middle CTA, cycles from its first stamp:")
t0 = d[0]
for k in range(0, len(d) - 1, 2):
    nxt = d[k + 2] if k + 2 < len(d) else d[-1]
    print(f"  tile {k // 2:2d}: start {d[k] - t0:7d}  wait-for-data {d[k + 1] - d[k]:6d}  tile total {nxt - d[k]:6d}")
print(f"  end {d[-1] - t0}")
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print(f"production build, back to back (warm L2): {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
tot = 0.0
for _ in range(10):
    flush.fill_(1)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
print(f"production build, L2 flushed before each launch: {tot / 10 * 1e3:.1f} us per launch")
