#!/bin/bash
# final-code ncu evidence: convolution kernels (--set full) + launch list of the bench step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:last_conv_tap_kernel -s 1 -c 1 -f -o gpurun_out/r2_lastconv python tools/run_one.py > gpurun_out/j26_a.log 2>&1; echo "last conv rc=$?"
timeout 300 $NCU -k regex:conv3x3_pair_kernel -s 8 -c 1 -f -o gpurun_out/r2_pairconv python tools/run_one.py > gpurun_out/j26_b.log 2>&1; echo "pair rc=$?"
timeout 300 $NCU -k regex:conv3x3_tc_kernel -s 3 -c 3 -f -o gpurun_out/r2_conv64 python tools/run_one.py > gpurun_out/j26_c.log 2>&1; echo "conv64 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/j26_bench.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches_final.csv
