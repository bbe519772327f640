#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "fused_attention" --timeout 120 2>&1 | tail -3
for c in 120 60 90; do
  timeout 120 python tools/attn2_timing.py $c 4 --full > gpurun_out/j3_attn2_timing_$c.txt 2>&1
  head -3 gpurun_out/j3_attn2_timing_$c.txt
done
