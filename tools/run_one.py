"""One or two forwards at the BASELINE cfg2 shape, for ncu captures."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
m = helpers.make_module(8, 4, "bf16").cuda().eval()
x = torch.rand(176, 1, 40, 32, device="cuda")
with torch.no_grad():
    for _ in range(n):
        y = m(x)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
