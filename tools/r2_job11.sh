#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_headmode.py tests/test_gpu_train_tc.py -m gpu -q --timeout 600 2>&1 | tail -6
