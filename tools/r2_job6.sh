#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stl_attn2 -s 3 -c 1 -f -o gpurun_out/r2_attn2_120 python tools/attn2_timing.py 120 4 > gpurun_out/j6_ncu.log 2>&1
tail -3 gpurun_out/j6_ncu.log
ls -la gpurun_out/*.ncu-rep
