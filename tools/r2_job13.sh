#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/j13_bench_n2.json 2> gpurun_out/j13_bench_n2.err
echo "rc=$?"
tail -c 1500 gpurun_out/j13_bench_n2.err
head -c 600 gpurun_out/j13_bench_n2.json
