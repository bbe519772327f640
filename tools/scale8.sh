T() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500+$1)) "${@:2}" 2>&1 | grep '^{' ; }
echo "== bench.py N=1,2,4,8"; python bench.py --no-cpu-baseline 2>/dev/null | grep '^{' | cut -c1-260
for n in 2 4 8; do T $n bench.py --gpus $n --no-cpu-baseline | cut -c1-260; done
echo "== cfg3 volume (strong)"; python tools/volume_bench.py; for n in 2 4 8; do T $n tools/volume_bench.py; done
echo "== cfg3 8 volumes"; python tools/volume_bench.py --volumes 8; T 8 tools/volume_bench.py --volumes 8
echo "== cfg4 train bf16 graph"; python tools/train_bench.py --precision bf16 --mode graph; for n in 2 4 8; do T $n tools/train_bench.py --precision bf16 --mode graph; done
echo "== cfg4 train bf16 eager torch DDP"; T 8 tools/train_bench.py --precision bf16 --mode eager --ddp torch
