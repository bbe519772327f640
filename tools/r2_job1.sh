#!/bin/bash
# round-2 job 1: headline parity tests, phase timelines of the attention kernel, a short bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/j1_smi.txt 2>&1
nproc >> gpurun_out/j1_smi.txt
timeout 900 python -m pytest tests/test_gpu_headline.py tests/test_gpu_tc.py "tests/test_gpu_network.py" -m gpu -q -x --timeout 600 > gpurun_out/j1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j1_pytest.log
for c in 60 90 120; do
  RDST_TIMING_ABS=1 timeout 120 python tools/attn_timing.py $c 4 > gpurun_out/j1_attn_timing_$c.txt 2>&1
done
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/j1_bench.json 2> gpurun_out/j1_bench.err
echo "bench rc=$?" >> gpurun_out/j1_bench.err
tail -3 gpurun_out/j1_pytest.log; cat gpurun_out/j1_bench.json
