#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_network.py -m gpu -q -x -k "graph" --timeout 200 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/j20_bench.json 2> gpurun_out/j20_bench.err; echo "rc=$?"; tail -3 gpurun_out/j20_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/j20_bench.json"))
for k in ("value","ms_per_step","e2e","roofline","rdst_e_cfg3","rdst_e_cfg3_strong","train_cfg4","parity_max_abs","clocks","cpu_baseline","gpu_launches"):
    v=d.get(k)
    if k=="roofline": v={kk:v[kk] for kk in ("achieved","frac","avg_launch_us","share_of_step","traffic")}
    print(k, v)
PY
