"""Where a training step's time goes (cfg4 shape 32x1x24x24, 1 GPU): wall time of forward / backward / optimizer,
and the per-ABI-call device time (CUDA events around every launch) summed over one step."""
import collections
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers  # noqa: E402
from rdst_b200 import _lib  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
torch.manual_seed(0)
m = helpers.make_module(8, 4, prec).cuda().train()
opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-4, betas=(0.9, 0.99), eps=1e-8, fused=True)
x = torch.rand(32, 1, 24, 24, device="cuda")
y = torch.rand(32, 1, 96, 96, device="cuda")


def step(t=None):
    def mark(k):
        if t is not None:
            torch.cuda.synchronize()
            t[k] = time.perf_counter()
    mark("t0")
    opt.zero_grad(set_to_none=True)
    out = m(x)
    loss = torch.nn.functional.l1_loss(out, y)
    mark("fwd")
    loss.backward()
    mark("bwd")
    opt.step()
    mark("opt")


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print(f"step (no instrumentation): {(time.perf_counter() - t0) / 5 * 1e3:.2f} ms")
t = {}
step(t)
print(f"forward {1e3 * (t['fwd'] - t['t0']):.2f} ms, backward {1e3 * (t['bwd'] - t['fwd']):.2f} ms, "
      f"optimizer {1e3 * (t['opt'] - t['bwd']):.2f} ms (each followed by a sync)")

orig = _lib.call
evs = []


def timed(name, *a):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    orig(name, *a)
    e1.record()
    if name == "rdst_gemm_tc":   # (x, ldx, w, ldw, w_mn, bias, resid, ldr, aux, lda, y, ldy, T, K, N, a_op, ln, scale, conv, ...)
        name += (f" T={a[12]} K={a[13]} N={a[14]}" + (" wT" if a[4] else "") + ("", " ln", " gelu")[a[15]] +
                 (" *gelu'" if a[8] is not None else "") + (" conv" if a[18] else ""))
    elif name in ("rdst_gemm_tn_tc", "rdst_gemm_tn_acc"):   # (dy, ldy, x, ldx, dw, db, T, N, K, conv, B, H, W, Cin[, x_op, creal])
        name += f" T={a[6]} N={a[7]} K={a[8]}" + (" conv" if a[9] else "") + (("", " ln", " gelu")[a[14]] if len(a) > 16 else "")
    evs.append((name, e0, e1))


_lib.call = timed
step()
torch.cuda.synchronize()
_lib.call = orig
agg = collections.defaultdict(lambda: [0, 0.0])
for k, a, b in evs:
    agg[k][0] += 1
    agg[k][1] += a.elapsed_time(b)
tot = sum(v for _, v in agg.values())
print(f"{'call':48s} {'n/step':>7s} {'us/call':>9s} {'ms/step':>8s} {'share':>6s}")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:48s} {n:7d} {v / n * 1e3:9.1f} {v:8.3f} {100 * v / tot:5.1f}%")
print(f"{'TOTAL device time in ABI calls':48s} {len(evs):7d} {'':9s} {tot:8.3f}")

# host-only cost: time to enqueue (no sync) one step
torch.cuda.synchronize()
t0 = time.perf_counter()
step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"host enqueue time of one step: {1e3 * (t1 - t0):.2f} ms")

# every CUDA kernel of one step (torch ops included), from the profiler: who owns the rest of the device time?
from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[2])
tot_n, tot_t = sum(r[1] for r in rows), sum(r[2] for r in rows)
ours = [r for r in rows if "rdst" in r[0]]
print(f"profiler: {tot_n} kernels, {tot_t / 1e3:.2f} ms device time; librdst kernels: {sum(r[1] for r in ours)} launches, "
      f"{sum(r[2] for r in ours) / 1e3:.2f} ms; torch kernels: {tot_n - sum(r[1] for r in ours)} launches, "
      f"{(tot_t - sum(r[2] for r in ours)) / 1e3:.2f} ms")
for k, n, t in rows[:14]:
    print(f"  {k[:90]:90s} {n:6d} {t / 1e3:8.3f} ms")
