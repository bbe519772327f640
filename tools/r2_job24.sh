#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for lib in $(ls rdst_b200/lib/ | grep probe_ | sed 's/.so//'); do
  RDST_B200_LIB=$PWD/rdst_b200/lib/$lib.so timeout 120 python tools/mlp2_timing.py 120 --full > gpurun_out/j24_$lib.txt 2>&1; sed -n 2p gpurun_out/j24_$lib.txt
done
