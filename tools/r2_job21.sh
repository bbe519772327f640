#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4 2 1; do
if [ $n -eq 1 ]; then
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/j21_bench_n$n.json 2> gpurun_out/j21_bench_n$n.err
else
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/j21_bench_n$n.json 2> gpurun_out/j21_bench_n$n.err
fi
echo "n=$n rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/j21_bench_n$n.json"))
print({k:(d[k] if not isinstance(d[k],dict) else {kk:d[k][kk] for kk in d[k] if kk in ("value","ms_per_step","ms_per_volume","frac")}) for k in ("n_gpus","value","ms_per_step","e2e","roofline","roofline_hbm_kernel","rdst_e_cfg3","rdst_e_cfg3_strong","train_cfg4")})
PY
done
