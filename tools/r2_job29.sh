#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -k "attention" --timeout 120 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_network.py tests/test_gpu_headline.py tests/test_swinir.py tests/test_headmode.py -m gpu -q --timeout 300 2>&1 | tail -4
for c in 120 90 60; do timeout 120 python tools/attn2_timing.py $c 0 2>&1 | sed -n 2,3p; done
timeout 100 python tools/bf16_error.py 2>&1 | tail -5
