#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "attention" --timeout 120 2>&1 | tail -6
for c in 60 90 120; do
  timeout 120 python tools/attn2_timing.py $c 0 --full > gpurun_out/j16_attn2_$c.txt 2>&1; head -3 gpurun_out/j16_attn2_$c.txt
done
