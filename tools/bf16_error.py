"""bf16-mode error of the drop-in vs the reference goldens (max abs, PSNR delta) -- prints, no asserts."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers
import rdst_oracle as O
for name in helpers.CASES:
    c = helpers.load_case(name)
    m = helpers.make_module(c["blocks"], c["scale"], "bf16").cuda().eval()
    m.load_state_dict(c["sd"], strict=True)
    with torch.no_grad():
        y = m(c["x"].cuda()).cpu()
    ref = torch.from_numpy(c["g"]["y"])
    tgt = helpers.realistic_target(ref)                  # reference output + noise at ~33 dB: the target of the PSNR tests
    print(f"{name:18s} max|err| {float((y-ref).abs().max()):.3e}  mean|err| {float((y-ref).abs().mean()):.3e}  dPSNR {O.psnr(y,tgt)-O.psnr(ref,tgt):+.5f} dB")
