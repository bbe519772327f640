"""Per-slice inference latency (the reference tester's access pattern): eager launches vs CUDA-graph replay."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers
from rdst_b200.infer import GraphedRDST
m = helpers.make_module(8, 4, "bf16").cuda().eval()
gm = GraphedRDST(m)
for B in (1, 8):
    x = torch.rand(B, 1, 40, 32, device="cuda")
    for name, fn in (("eager", lambda: m(x)), ("graph", lambda: gm(x))):
        with torch.no_grad():
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"B={B} {name}: {ms:.3f} ms/forward -> {B * 160 * 128 / ms / 1e3:.2f} HR Mpix/s")
