#!/bin/bash
# final-code ncu captures for profiles/: attention v2 (C=120, C=60), MLP v2 (C=120 plain + tail)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stl_attn2 -s 3 -c 1 -f -o gpurun_out/r2_attn2_120 python tools/attn2_timing.py 120 4 > gpurun_out/j31_a120.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stl_attn2 -s 3 -c 1 -f -o gpurun_out/r2_attn2_60 python tools/attn2_timing.py 60 4 > gpurun_out/j31_a60.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stl_mlp2 -s 3 -c 1 -f -o gpurun_out/r2_mlp2_120 python tools/mlp2_timing.py 120 > gpurun_out/j31_m120.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stl_mlp2 -s 3 -c 1 -f -o gpurun_out/r2_mlp2_tail_120 python tools/mlp2_timing.py 120 --tail > gpurun_out/j31_mt120.log 2>&1; echo rc=$?
