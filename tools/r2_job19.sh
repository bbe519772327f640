#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 120 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "last_conv" --timeout 60 2>&1 | tail -12
timeout 100 python tools/lastconv_timing.py
