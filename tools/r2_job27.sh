#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'stl_|conv|head_kernel|layernorm|last_conv' --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/j27_bench_under_ncu.log 2>&1; echo "rc=$?"
