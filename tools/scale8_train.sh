# final-code training scaling on one 8-GPU box: N=1 and N=8 (graph, bf16 GEMMs + attention, BucketedAllReduce)
python tools/train_bench.py --precision bf16 --mode graph --steps 20 2>&1 | grep '^{'
python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-8} --master-addr 127.0.0.1 --master-port 29531 tools/train_bench.py --precision bf16 --mode graph --steps 20 2>&1 | grep '^{'
