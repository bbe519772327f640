"""Per-ABI-call time of one bf16 forward at the cfg2 shape, measured with CUDA events around every launch (real clocks)."""
import os, sys, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers
from rdst_b200 import executor as ex
PREC = sys.argv[1] if len(sys.argv) > 1 else "bf16"
m = helpers.make_module(8, 4, PREC).cuda().eval()
x = torch.rand(176, 1, 40, 32, device="cuda")
with torch.no_grad():
    for _ in range(3): m(x)
torch.cuda.synchronize()
orig = ex.call
evs = []
def timed(name, *a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); orig(name, *a); e1.record()
    key = name
    if name == "rdst_stl_attn_fwd_bf16": key += f" C={a[12]}"
    elif name.startswith("rdst_stl_mlp"): key += f" C={a[-3]}"
    elif name == "rdst_conv3x3_fwd_bf16_tc": key += f" {a[11]}->{a[12]} @{a[9]}x{a[10]}"
    evs.append((key, e0, e1))
ex.call = timed
N = 5
with torch.no_grad():
    for _ in range(N): m(x)
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for k, a, b in evs:
    agg[k][0] += 1; agg[k][1] += a.elapsed_time(b)
tot = sum(v for _, v in agg.values())
print(f"{'call':58s} {'n/fwd':>6s} {'us/call':>9s} {'ms/fwd':>8s} {'share':>6s}")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:58s} {n / N:6.0f} {v / n * 1e3:9.1f} {v / N:8.3f} {100 * v / tot:5.1f}%")
print(f"{'TOTAL (sum of per-call event times)':58s} {len(evs) / N:6.0f} {'':9s} {tot / N:8.3f}")
