#!/bin/bash
# full GPU suite + bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/j7_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j7_pytest.log
tail -4 gpurun_out/j7_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/j7_bench.json 2> gpurun_out/j7_bench.err
cat gpurun_out/j7_bench.json
