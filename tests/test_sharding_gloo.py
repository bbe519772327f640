"""world_size-2 gloo test (CPU) of the slice-sharded inference driver: every rank runs its contiguous range through
the drop-in module (kernels replaced by their contract restatements), results are gathered and must equal the
single-process result bit for bit -- the N>1 path has no data-path collective, only the final gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_slices, ret):
    import sys
    for p in (helpers.ROOT, os.path.join(helpers.ROOT, "oracle"), os.path.join(helpers.ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from abi_emulator import emulated_abi
    from rdst_b200 import infer
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    c = helpers.load_case("e2blk_x4_8x8")
    m = helpers.make_module(c["blocks"], c["scale"])
    m.load_state_dict(c["sd"])
    vol = torch.rand(n_slices, 1, 8, 16, generator=torch.Generator().manual_seed(77))
    b, e = infer.shard_range(n_slices, world, rank)
    with emulated_abi(), torch.no_grad():
        mine = m._exec._forward_impl(vol[b:e])
    sizes = [infer.shard_range(n_slices, world, r) for r in range(world)]
    parts = [torch.empty(se - sb, 1, 32, 64) for sb, se in sizes]
    dist.all_gather(parts, mine) if len({p.shape for p in parts}) == 1 else None
    if len({p.shape for p in parts}) != 1:            # ragged split: exchange with padding-free point-to-point
        for r in range(world):
            if r == rank:
                parts[r] = mine
            dist.broadcast(parts[r], src=r)
    full = torch.cat(parts)
    if rank == 0:
        with emulated_abi(), torch.no_grad():
            ref = m._exec._forward_impl(vol)
        ret["equal"] = bool(torch.equal(full, ref))
        ret["shape"] = tuple(full.shape)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_slices", [6, 7])
def test_two_rank_sharded_inference_matches_single_process(n_slices):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, n_slices, ret), nprocs=2, join=True)
    assert ret["equal"] and ret["shape"] == (n_slices, 1, 32, 64)


def test_shard_range_partitions_exactly():
    from rdst_b200.infer import shard_range
    for n in (0, 1, 7, 176, 177):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
