"""Eager launches vs CUDA-graph replay (rdst_b200.infer.GraphedRDST) of one forward: a single OASIS slice (the reference
tester's access pattern) and the 176-slice volume.  Measured: 2.50 -> 1.62 ms per slice, 8.23 -> 8.15 ms per volume."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, _p)
import helpers
from rdst_b200.infer import GraphedRDST
m = helpers.make_module(8, 4, "bf16").cuda().eval()
gm = GraphedRDST(m)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for B in (1, 176):
    x = torch.rand(B, 1, 40, 32, device="cuda")
    for name, fn in (("eager", lambda: m(x)), ("graph", lambda: gm(x))):
        with torch.no_grad():
            for _ in range(3): fn()
            tot = 0
            for _ in range(10):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
        print(f"B={B} {name}: {tot/10:.3f} ms")
