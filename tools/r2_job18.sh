#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "fused_mlp" --timeout 120 2>&1 | tail -4
for c in 120 90 60; do
  timeout 120 python tools/mlp2_timing.py $c --full > gpurun_out/j18_mlp2_$c.txt 2>&1; head -3 gpurun_out/j18_mlp2_$c.txt
  timeout 120 python tools/mlp2_timing.py $c --tail --full > gpurun_out/j18_mlp2_tail_$c.txt 2>&1; head -3 gpurun_out/j18_mlp2_tail_$c.txt
done
