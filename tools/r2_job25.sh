#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 200 python -m pytest tests/test_gpu_tc.py -m gpu -q -k "conv3x3" --timeout 60 2>&1 | tail -3
timeout 400 python -m pytest tests/test_gpu_network.py tests/test_gpu_headline.py -m gpu -q -x --timeout 200 2>&1 | tail -3
timeout 200 python tools/kernel_breakdown.py 2>&1 | grep -E "conv3x3|last_conv|TOTAL"
