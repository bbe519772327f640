#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -k "fused_mlp" --timeout 120 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_network.py tests/test_gpu_headline.py tests/test_swinir.py tests/test_headmode.py tests/test_3conv.py tests/test_rdstn.py tests/test_estsr.py -m gpu -q --timeout 300 2>&1 | tail -4
for c in 120 90 60; do timeout 120 python tools/mlp2_timing.py $c --tail 2>&1 | sed -n 1,3p; done
