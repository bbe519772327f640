"""Per-MMA cost of short tcgen05 accumulate chains (rdst_umma_bench): dependent chain vs independent accumulators."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdst_b200 import _lib
out = torch.zeros(2, dtype=torch.int64, device="cuda")
def run(N, chains, count, a_tmem, masked):
    best = None
    for _ in range(3):
        _lib.call("rdst_umma_bench", N, chains, count, a_tmem, masked, _lib.ptr(out), _lib.stream_ptr())
        torch.cuda.synchronize()
        t = out.cpu().tolist()
        best = t if best is None or t[0] < best[0] else best
    return best
print("N chains count A masked | total cyc, issue cyc | cyc/MMA (floor = N/2)")
combos = []
for N in (16, 32, 64, 128):
    combos += [(N, 1, 1, 0), (N, 2, 1, 0), (N, 1, 1, 1), (N, 2, 1, 1), (N, 1, 0, 0), (N, 2, 0, 0)]
combos += [(32, 4, 1, 0), (64, 4, 1, 0), (32, 4, 1, 1), (64, 4, 1, 1), (256, 1, 1, 0), (256, 1, 0, 0), (192, 1, 1, 0), (192, 2, 1, 0)]
for N, chains, a_tmem, masked in combos:
    for count in (16, 128):
        tot, iss = run(N, chains, count, a_tmem, masked)
        print(f"{N:4d} {chains:2d} {count:3d} {'TS' if a_tmem else 'SS'} {masked} | {tot:6d} {iss:6d} | {tot / count:7.1f}  (floor {N / 2:.0f})")
