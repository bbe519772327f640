"""Phase-level cycle timeline of CTA 0 of the fused MLP kernel (uses rdst_debug_mlp_timing)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdst_b200 import _lib, packing
c = int(sys.argv[1]) if len(sys.argv) > 1 else 120
T = 176 * 40 * 32
cp, hp = packing.padded_width(c), packing.hidden_width(2 * c)
pos = packing.channel_positions(c)
g = torch.Generator().manual_seed(0)
x = torch.zeros(T, cp); x[:, pos] = torch.randn(T, c, generator=g); x = x.to(torch.bfloat16).cuda()
w1 = torch.zeros(hp, cp); w1[:2 * c, pos] = torch.randn(2 * c, c, generator=g) * 0.08
w2 = torch.zeros(cp, hp); w2[pos, :2 * c] = torch.randn(c, 2 * c, generator=g) * 0.08
d = [packing.fc1_image(w1).cuda(), packing.fc2_image(w2).cuda(), torch.zeros(hp).cuda(), torch.zeros(cp).cuda()]
y = torch.empty_like(x)
dbg = torch.zeros(128, dtype=torch.int64, device="cuda")
run = lambda: _lib.call("rdst_stl_mlp_fwd_bf16", _lib.ptr(x), cp, _lib.ptr(y), cp, _lib.ptr(d[0]), _lib.ptr(d[1]), _lib.ptr(d[2]),
                        _lib.ptr(d[3]), T, c, 0, _lib.stream_ptr())
run(); torch.cuda.synchronize()
_lib.call("rdst_debug_mlp_timing", _lib.ptr(dbg)); run(); torch.cuda.synchronize(); _lib.call("rdst_debug_mlp_timing", None)
nchk = (hp + 63) // 64
names = ["tile start", "P1a done", "P1b done"]
for k in range(nchk):
    names += [f"before fc1[{k}] wait", f"fc1[{k}] ready", f"GELU[{k}] done"]
names += ["before fc2 wait", "fc2 ready", "P5 done", "tile done"]
t = dbg.cpu().tolist()
for half in range(2):
    print(f"--- warpgroup {half} (C={c})")
    for tile in range(2):
        base = t[half * 64 + tile * len(names)]; prev = base
        for k, nm in enumerate(names):
            v = t[half * 64 + tile * len(names) + k]
            print(f"  {nm:22s} +{v - prev:6d}  (t={v - base:6d})"); prev = v
