#!/bin/bash
# full GPU suite, smoke, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/j9_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j9_pytest.log
tail -6 gpurun_out/j9_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j9_smoke.log 2>&1; echo "smoke rc=$?"; tail -8 gpurun_out/j9_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/j9_bench.json 2> gpurun_out/j9_bench.err
python -c "
import json;d=json.load(open('gpurun_out/j9_bench.json'));print({k:d[k] for k in ('value','ms_per_step','parity_max_abs','gpu_launches')}, d['roofline']['frac'], d['roofline']['avg_launch_us'], d['e2e']['value'], d.get('rdst_e_cfg3',{}).get('ms_per_step'), d.get('train_cfg4',{}).get('ms_per_step'))"
