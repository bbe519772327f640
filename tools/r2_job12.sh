#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
python - <<'PY'
import sys, os
for p in (".", "oracle", "tests"): sys.path.insert(0, os.path.abspath(p))
import torch, helpers, rdst_oracle as O
for name in helpers.HEADMODE_CASES + helpers.C3_CASES:
    c = helpers.load_3conv_case(name)
    ref = torch.from_numpy(c["g"]["y"])
    mk = helpers.make_headmode if "head" in name else helpers.make_3conv
    m = mk(c, "bf16").cuda().eval(); m.load_state_dict(c["sd"])
    with torch.no_grad(): y = m(c["x"].cuda()).cpu()
    tgt = helpers.realistic_target(ref)
    e = y - ref
    print(name, "max", float(e.abs().max()), "rms", float(e.pow(2).mean().sqrt()), "range", float(ref.min()), float(ref.max()), "dPSNR", O.psnr(y, tgt) - O.psnr(ref, tgt))
PY
