"""Data-parallel training step of RDST-E1 x4 (BASELINE cfg4 shape: 32 x 1x24x24 -> 32 x 1x96x96 per GPU, L1 loss,
Adam lr 1e-4 betas (0.9, 0.99)) with torch DDP over NCCL.  Launch with torchrun for N > 1.
Prints HR Mpix/s = ranks * 32 * 96^2 / step time and checks that all ranks hold identical weights afterwards."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    torch.manual_seed(0)
    m = helpers.make_module(blocks, 4, "fp32").cuda().train()
    model = torch.nn.parallel.DistributedDataParallel(m, device_ids=[local]) if world > 1 else m
    opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-4, betas=(0.9, 0.99), eps=1e-8)
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    x = torch.rand(32, 1, 24, 24, device="cuda", generator=g)
    y = torch.rand(32, 1, 96, 96, device="cuda", generator=g)
    losses = []
    for it in range(steps + 2):
        if it == 2:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.l1_loss(model(x), y)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = (time.perf_counter() - t0) / steps
    # replicas must stay bit-identical (same init, all-reduced gradients)
    chk = torch.stack([p.detach().double().sum() for p in m.parameters()]).sum()
    if world > 1:
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(lo == hi)
    else:
        same = True
    if rank == 0:
        print(f"train step: world={world} blocks={blocks} {dt * 1e3:.1f} ms/step -> "
              f"{world * 32 * 96 * 96 / dt / 1e6:.3f} HR Mpix/s; loss {losses[0]:.4f} -> {losses[-1]:.4f}; "
              f"replicas identical: {same}")
        assert same and losses[-1] < losses[0]
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
