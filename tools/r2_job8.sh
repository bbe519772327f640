#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/j8_swinir_psnr.txt 2>&1
import sys, os
for p in (".", "oracle", "tests"): sys.path.insert(0, os.path.abspath(p))
import torch, helpers, rdst_oracle as O
from rdst_b200 import _lib
for name in helpers.SWINIR_CASES:
    c = helpers.load_swinir_case(name)
    ref = torch.from_numpy(c["g"]["y"])
    for variant in (1, 2):
        _lib.call("rdst_debug_attn_variant", variant)
        m = helpers.make_swinir(c, "bf16").cuda().eval(); m.load_state_dict(c["sd"])
        with torch.no_grad(): y = m(c["x"].cuda()).cpu()
        tgt = helpers.realistic_target(ref)
        e = (y - ref)
        print(name, "variant", variant, "max", float(e.abs().max()), "rms", float(e.pow(2).mean().sqrt()), "ref range", float(ref.min()), float(ref.max()),
              "dPSNR", O.psnr(y, tgt) - O.psnr(ref, tgt))
PY
cat gpurun_out/j8_swinir_psnr.txt
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_headline.py tests/test_gpu_network.py -m gpu -q -x --timeout 600 2>&1 | tail -3
for c in 60 90 120; do timeout 120 python tools/attn2_timing.py $c 4 --full > gpurun_out/j3_attn2_timing_$c.txt 2>&1; head -2 gpurun_out/j3_attn2_timing_$c.txt | tail -1; done
