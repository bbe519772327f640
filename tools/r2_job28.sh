#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py tests/test_gpu_backward.py tests/test_3conv.py tests/test_rdstn.py -m gpu -q -x --timeout 300 2>&1 | tail -4
timeout 300 python tools/kernel_breakdown.py fp32 2>&1 | tail -8
