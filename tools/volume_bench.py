"""BASELINE cfg3: RDST-E (4 RDSTBs) x4, a synthetic OASIS-shaped volume (176 LR slices of 40x32) super-resolved with the
slice axis sharded over the ranks (rdst_b200.infer.super_resolve_volume_sharded; no data-path collective).
Strong scaling: total work fixed, each rank takes a contiguous slice range.  Launch with torchrun for N > 1.

    python tools/volume_bench.py [--volumes V] [--blocks 4] [--steps K]

Prints one JSON line: HR Mpix/s of the whole job = V*176*160*128 / max-over-ranks device time (inputs resident in HBM),
and the same end to end from pinned host memory (H2D + forward + D2H per rank).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import helpers  # noqa: E402
from rdst_b200 import infer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--volumes", type=int, default=1)
    ap.add_argument("--blocks", type=int, default=4)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(0)
    m = helpers.make_module(a.blocks, 4, "bf16").cuda().eval()
    n = 176 * a.volumes
    vol = torch.rand(n, 1, 40, 32, generator=torch.Generator().manual_seed(1)).pin_memory()      # same volume on every rank
    b, e = infer.shard_range(n, world, rank)
    dev_in = vol[b:e].cuda()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        tot = 0.0
        for _ in range(a.steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        t = torch.tensor([tot / a.steps], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    ms_res = timed(lambda: infer.super_resolve_slices(m, dev_in, batch_size=176))
    ms_e2e = timed(lambda: infer.super_resolve_volume_sharded(m, vol, rank, world, batch_size=176))
    mpix = n * 160 * 128 / 1e6
    if rank == 0:
        print(json.dumps({"what": f"RDST-E ({a.blocks} RDSTB) x4 volume inference, {n} LR slices 40x32 sharded over {world} GPU(s)",
                          "world": world, "slices_per_rank": e - b, "scaling": "strong", "ms_resident": round(ms_res, 3),
                          "hr_mpix_per_s_resident": round(mpix / ms_res * 1e3, 1), "ms_e2e": round(ms_e2e, 3),
                          "hr_mpix_per_s_e2e": round(mpix / ms_e2e * 1e3, 1)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
