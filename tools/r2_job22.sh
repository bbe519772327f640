#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/train_bench.py --precision bf16 --mode graph --steps 20 --check 2>&1 | grep '^{' | cut -c1-330,500-800
RDST_DDP_BUCKETS=3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tools/train_bench.py --precision bf16 --mode graph --steps 20 2>&1 | grep '^{' | cut -c1-330
