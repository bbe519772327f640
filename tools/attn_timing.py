"""Phase-level cycle timeline of CTA 0 of the fused attention kernel (uses rdst_debug_attn_timing)."""
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
run(); torch.cuda.synchronize()
_lib.call("rdst_debug_attn_timing", _lib.ptr(dbg))
run(); torch.cuda.synchronize()
_lib.call("rdst_debug_attn_timing", None)
d = dbg.cpu().tolist()
names = ["tile start", "tile landed", "stats written", "P1a done", "P1b loaded", "P1b stored", "P1b done"]
for i in range(3):
    names += [f"h{i} before qkv wait", f"h{i} drained", f"h{i} S ready", f"h{i} bias+max", f"h{i} softmax done"]
names += ["heads issued", "O epi+proj issued", "proj done", "proj loaded", "y written", "y staged", "tile done"]
for wg in range(2):
    t = d[wg * 64: wg * 64 + 64]
    print(f"--- warpgroup {wg} (C={c}, shift={shift}); first two tiles")
    n = len(names)
    for tile in range(2):
        base = t[tile * n]
        prev = base
        for k, nm in enumerate(names):
            v = t[tile * n + k]
            print(f"  {nm:22s} +{v - prev:6d}  (t={v - base:6d})")
            prev = v
if os.environ.get("RDST_TIMING_ABS"):
    t0 = d[0]
    print("--- absolute cycles (both stamping threads, relative to slot 0 tile 0 start)")
    for tile in range(2):
        for k, nm in enumerate(names):
            print(f"  tile{tile} {nm:22s} slot0 {d[tile * n + k] - t0:7d}   slot1 {d[64 + tile * n + k] - t0:7d}")
