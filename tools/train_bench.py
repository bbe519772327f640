"""Data-parallel training step of RDST-E1 x4 (BASELINE cfg4 shape: 32 x 1x24x24 -> 32 x 1x96x96 per GPU, L1 loss,
Adam lr 1e-4 betas (0.9, 0.99)).  Launch with torchrun for N > 1.

    python tools/train_bench.py [--steps K] [--blocks 8] [--mode eager|graph] [--ddp auto|torch|bucket]
                                [--precision fp32|bf16] [--check]

mode=graph   : rdst_b200.train.GraphedTrainStep (whole step = one CUDA graph; N > 1 uses ddp=bucket)
ddp=torch    : torch DistributedDataParallel (eager only); ddp=bucket: rdst_b200.ddp.BucketedAllReduce
Prints one JSON line: HR Mpix/s = ranks * 32 * 96^2 / step time (device time, max over ranks), and checks that all
ranks hold identical weights afterwards and that the loss went down.  --check also runs the same steps eagerly from
the same initial weights and reports the largest parameter difference (graph replay == eager).
"""
import argparse
import copy
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--blocks", type=int, default=8)
    ap.add_argument("--mode", default="graph", choices=["eager", "graph"])
    ap.add_argument("--ddp", default="auto", choices=["auto", "torch", "bucket"])
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--model", default="rdst", choices=["rdst", "rdstn", "swinir"])
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ddp = a.ddp if a.ddp != "auto" else ("bucket" if a.mode == "graph" else "torch")
    from rdst_b200 import ddp as rddp, train as rtrain

    torch.manual_seed(0)
    def build():
        if a.model == "swinir":      # the ini's [SwinIR] section: 4 RSTBs x 6 blocks, constructor resolution 8 (no shifted blocks)
            return helpers.make_swinir(dict(img_size=8, depths=[6] * 4, upscale=4), a.precision)
        if a.model == "rdstn":
            return helpers.make_rdstn(dict(blocks=a.blocks, scale=4), a.precision)
        return helpers.make_module(a.blocks, 4, a.precision)

    m = build().cuda().train()
    m0 = copy.deepcopy(m.state_dict())
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    x = torch.rand(a.batch, 1, 24, 24, device="cuda", generator=g)
    y = torch.rand(a.batch, 1, 96, 96, device="cuda", generator=g)

    def make_opt(model, capturable):
        # the reference's optimiser (utils/optim.py: Adam, ini :130-135); fused=True = one multi-tensor kernel
        return torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-4, betas=(0.9, 0.99),
                                eps=1e-8, capturable=capturable, fused=True)

    def run(model, mode, steps, timed):
        opt = make_opt(model, mode == "graph")
        nb = int(os.environ.get("RDST_DDP_BUCKETS", "0"))      # experiment: 1 = one all-reduce at the end, 3 = tail+body.4-7 | body.0-3 | head
        def coarse(name):
            k = rddp.rdst_link_of(name)
            if nb == 1:
                return "all"
            if k.startswith("body."):
                return "hi" if int(k.split(".")[1]) >= 4 else "lo"
            return "hi" if k == "tail" else "lo2"
        reducer = (rddp.BucketedAllReduce(model, bucket_of=coarse) if nb else rddp.BucketedAllReduce(model)) if ((world > 1 and ddp == "bucket") or os.environ.get("RDST_FORCE_REDUCER")) else None
        wrapped = model
        if world > 1 and ddp == "torch":
            wrapped = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], bucket_cap_mb=4)
        if mode == "graph":
            stepper = rtrain.GraphedTrainStep(model, opt, x, y, reducer=reducer)
            step = lambda: stepper(x, y)
        else:
            def step():
                if reducer is not None:
                    reducer.begin_step()
                else:
                    opt.zero_grad(set_to_none=True)
                loss = torch.nn.functional.l1_loss(wrapped(x), y)
                loss.backward()
                if reducer is not None:
                    reducer.finish()
                opt.step()
                return loss.detach()
        losses = []
        warm = 2 if timed else 0
        for _ in range(warm):
            losses.append(step().clone())
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(steps):
            losses.append(step().clone())
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if reducer is not None:
            order = list(reducer.launch_order)
            reducer.remove()
        else:
            order = None
        return float(ms), [float(l) for l in losses], order

    ms, losses, order = run(m, a.mode, a.steps, True)
    chk = torch.stack([p.detach().double().sum() for p in m.parameters()]).sum()
    same = True
    if world > 1:
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(lo == hi)
    line = {"what": {"rdst": "RDST-E1", "rdstn": "RDSTSR_N", "swinir": "SwinIR-lite"}[a.model] + " x4 training step (L1, Adam)",
            "world": world, "blocks": a.blocks, "mode": a.mode,
            "ddp": ddp if world > 1 else None, "precision": a.precision, "batch_per_gpu": a.batch,
            "ms_per_step": round(ms, 3), "hr_mpix_per_s": round(world * a.batch * 96 * 96 / ms / 1e3, 3),
            "loss_first": round(losses[0], 5), "loss_last": round(losses[-1], 5), "replicas_identical": same,
            "bucket_launch_order": order}
    if a.check:
        m2 = build().cuda().train()
        m2.load_state_dict(m0)
        _, losses2, _ = run(m2, "eager", a.steps + 2, False)
        diff = max((p.detach() - q.detach()).abs().max().item() for p, q in zip(m.parameters(), m2.parameters()))
        line["max_param_diff_vs_eager"] = diff
        line["loss_last_eager"] = round(losses2[-1], 5)
    if rank == 0:
        print(json.dumps(line))
        assert same and losses[-1] < losses[0], line
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
