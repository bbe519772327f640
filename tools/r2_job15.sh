#!/bin/bash
# ncu captures for profiles/: attention v2 (C=120, C=60), MLP v2 (C=120 plain + tail); launch list of the bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stl_attn2 -s 3 -c 1 -f -o gpurun_out/r2_attn2_120 python tools/attn2_timing.py 120 4 > gpurun_out/j15_ncu_a120.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stl_attn2 -s 3 -c 1 -f -o gpurun_out/r2_attn2_60 python tools/attn2_timing.py 60 4 > gpurun_out/j15_ncu_a60.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stl_mlp2 -s 3 -c 1 -f -o gpurun_out/r2_mlp2_120 python tools/mlp2_timing.py 120 > gpurun_out/j15_ncu_m120.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stl_mlp2 -s 3 -c 1 -f -o gpurun_out/r2_mlp2_tail_120 python tools/mlp2_timing.py 120 --tail > gpurun_out/j15_ncu_mt120.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'stl_|conv|head_kernel|layernorm|last_conv' --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/j15_bench_under_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches.csv
