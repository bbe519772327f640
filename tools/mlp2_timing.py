"""Launch timing of the fused MLP kernels, lock-step (variant 1) vs warp-specialised (variant 2), plain and TAIL variants;
role timelines of CTA 0 (rdst_debug_mlp2_timing) with --full.      python tools/mlp2_timing.py C [--tail] [--full]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
from rdst_b200 import _lib, packing
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
c = int(argv[0]) if argv else 120
tail = "--tail" in sys.argv
T = 176 * 40 * 32
cp, hp, pos = packing.padded_width(c), packing.hidden_width(2 * c), packing.channel_positions(c)
g = torch.Generator().manual_seed(0)
x = torch.zeros(T, cp); x[:, pos] = torch.randn(T, c, generator=g)
x = x.to(torch.bfloat16).cuda()
w1 = torch.zeros(hp, cp); w1[:2 * c, pos] = torch.randn(2 * c, c, generator=g) * 0.08
w2 = torch.zeros(cp, hp); w2[pos, :2 * c] = torch.randn(c, 2 * c, generator=g) * 0.08
wt = torch.zeros(32, cp); wt[:30, pos] = torch.randn(30, c, generator=g) * 0.1
dev = [packing.fc1_image(w1).cuda(), packing.fc2_image(w2).cuda(), torch.zeros(hp).cuda(), torch.zeros(cp).cuda(),
       packing.kmajor_image(wt).cuda(), torch.zeros(32).cuda()]
y = torch.empty_like(x)
dense = torch.zeros(T, 160, dtype=torch.bfloat16, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run():
    if tail:
        _lib.call("rdst_stl_mlp_tail_fwd_bf16", _lib.ptr(x), cp, _lib.ptr(dev[0]), _lib.ptr(dev[1]), _lib.ptr(dev[2]), _lib.ptr(dev[3]),
                  _lib.ptr(dev[4]), _lib.ptr(dev[5]), _lib.ptr(dense[:, 96:]), 160, 0.5, T, c, 0, _lib.stream_ptr())
    else:
        _lib.call("rdst_stl_mlp_fwd_bf16", _lib.ptr(x), cp, _lib.ptr(y), cp, _lib.ptr(dev[0]), _lib.ptr(dev[1]), _lib.ptr(dev[2]),
                  _lib.ptr(dev[3]), T, c, 0, _lib.stream_ptr())
outs = {}
for variant in (1, 2):
    _lib.call("rdst_debug_mlp_variant", variant)
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
    outs[variant] = (dense if tail else y).clone()
    print(f"variant {variant}: C={c} tail={tail}: {sum(ts) / len(ts):8.2f} us/launch (min {min(ts):.2f}); {T // 128} tiles")
d = (outs[1].float() - outs[2].float()).abs()
print(f"variant 1 vs 2: max abs diff {d.max().item():.4e}, mean {d.mean().item():.4e}")
_lib.call("rdst_debug_mlp_variant", 2)
if "--full" in sys.argv:
    dbg = torch.zeros(1280, dtype=torch.int64, device="cuda")
    _lib.call("rdst_debug_mlp2_timing", _lib.ptr(dbg))
    run(); torch.cuda.synchronize()
    _lib.call("rdst_debug_mlp2_timing", None)
    dl = dbg.cpu().tolist()
    t0 = min(v for r in range(5) for v in dl[r * 256 + 1: r * 256 + 256] if v > 0)
    for r, name in enumerate(["A (LayerNorm)", "G0", "G1", "E (epilogue)", "I (fc1 issuer)"]):
        st = [v - t0 for v in dl[r * 256 + 1: r * 256 + 256] if v > 0]
        print(f"role {name} ({len(st)} stamps): first {st[:1]} last {st[-1:]} deltas {[b - a for a, b in zip(st, st[1:])][:48]}")
