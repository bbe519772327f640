#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for c in 120 60; do
  RDST_B200_LIB=$PWD/rdst_b200/lib/probe_FINE.so timeout 120 python tools/mlp2_timing.py $c --full > gpurun_out/j17_mlpfine_$c.txt 2>&1; sed -n 1,2p gpurun_out/j17_mlpfine_$c.txt
done
