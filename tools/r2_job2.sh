#!/bin/bash
# round-2 job 2: first run of the warp-specialised attention kernel (hang-guarded)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "fused_attention" --timeout 120 > gpurun_out/j2_pytest_attn.log 2>&1
rc=$?
echo "pytest attn rc=$rc" >> gpurun_out/j2_pytest_attn.log
tail -15 gpurun_out/j2_pytest_attn.log
if [ $rc -ne 0 ]; then
  # which variants / sizes fail: run each attention case alone
  timeout 400 python -m pytest tests/test_gpu_tc.py -m gpu -q -k "fused_attention" --timeout 60 > gpurun_out/j2_pytest_attn_all.log 2>&1
  tail -30 gpurun_out/j2_pytest_attn_all.log
fi
for c in 120 60 90; do
  timeout 120 python tools/attn2_timing.py $c 4 --full > gpurun_out/j2_attn2_timing_$c.txt 2>&1
  head -4 gpurun_out/j2_attn2_timing_$c.txt
done
if [ $rc -eq 0 ]; then
  timeout 600 python -m pytest tests/test_gpu_headline.py tests/test_gpu_network.py -m gpu -q -x --timeout 600 > gpurun_out/j2_pytest_net.log 2>&1
  echo "pytest net rc=$?" >> gpurun_out/j2_pytest_net.log
  tail -5 gpurun_out/j2_pytest_net.log
  timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/j2_bench.json 2> gpurun_out/j2_bench.err
  cat gpurun_out/j2_bench.json
fi
