"""Quick device-side timing of the module forward at the BASELINE cfg2 shape (176 x 1x40x32)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 176
for prec in sys.argv[2:] or ["fp32", "bf16"]:
    m = helpers.make_module(8, 4, prec).cuda().eval()
    x = torch.rand(B, 1, 40, 32, device="cuda")
    with torch.no_grad():
        for _ in range(2):
            y = m(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        n = 3
        e0.record()
        for _ in range(n):
            y = m(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{prec}: B={B} {ms:.2f} ms/forward -> {B * 160 * 128 / ms / 1e3:.2f} HR Mpix/s")
