"""Role timelines of CTA 0 of the warp-specialised attention kernel (rdst_debug_attn2_timing) + plain launch timing.

    python tools/attn2_timing.py C [shift] [B]
Prints, per role (A = LayerNorm/epilogue/TMA, B = drain/O-norm, C/D = softmax, M = MMA issue), the clock64 stamps of one
thread relative to the first stamp of the CTA, and the average launch duration of both kernel variants."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
from rdst_b200 import _lib, packing
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
c = int(argv[0]) if len(argv) > 0 else 120
shift = int(argv[1]) if len(argv) > 1 else 0
B = int(argv[2]) if len(argv) > 2 else 176
H, W = 40, 32
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run():
    _lib.call("rdst_stl_attn_fwd_bf16", _lib.ptr(x), cp, _lib.ptr(y), cp, _lib.ptr(pk["wqkv_img"]), _lib.ptr(pk["wproj_img"]),
              _lib.ptr(pk["bqkv_tc"]), _lib.ptr(bp), _lib.ptr(pk["table_tc"]), B, H, W, c, shift, _lib.stream_ptr())
outs = {}
for variant in (1, 2):
    _lib.call("rdst_debug_attn_variant", variant)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); run(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    outs[variant] = y.clone()
    ntiles = (B * (H // 8) * (W // 8) + 1) // 2
    per_cta = -(-ntiles // 148)
    print(f"variant {variant}: C={c} shift={shift} B={B}: {sum(ts) / len(ts):8.2f} us/launch (min {min(ts):.2f}); "
          f"{ntiles} tiles, {per_cta} per CTA -> {min(ts) * 1.965e3 / per_cta:.0f} cycles/tile @1.965GHz")
d = (outs[1].float() - outs[2].float()).abs()
print(f"variant 1 vs 2: max abs diff {d.max().item():.4e}, mean {d.mean().item():.4e}")
_lib.call("rdst_debug_attn_variant", 2)
dbg = torch.zeros(1280, dtype=torch.int64, device="cuda")
_lib.call("rdst_debug_attn2_timing", _lib.ptr(dbg))
run(); torch.cuda.synchronize()
_lib.call("rdst_debug_attn2_timing", None)
d = dbg.cpu().tolist()
roles = "ABCDM"
t0 = min(v for r in range(5) for v in d[r * 256 + 1: r * 256 + 256] if v > 0)
for r in range(5):
    st = [v - t0 for v in d[r * 256 + 1: r * 256 + 256] if v > 0]
    if "--full" in sys.argv:
        print(f"role {roles[r]} ({len(st)} stamps):", " ".join(str(v) for v in st[:200]))
    else:
        deltas = [b - a for a, b in zip(st, st[1:])]
        print(f"role {roles[r]} ({len(st)} stamps): first {st[:1]}, last {st[-1:]}; deltas[:60] = {deltas[:60]}")
